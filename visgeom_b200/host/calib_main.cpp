// vg_calib -- the reference's `calib` tool (test/calibration/generic_calibration.cpp:32-44) on the CUDA engine:
//   vg_calib problem.json [problem2.json ...]
// Every file adds its cameras, transformations and datasets; then one global solve, the report on stdout and
// image_error_<i>.txt per dataset.  Options (before the files): --device N, --precision N (digits printed,
// default 6 like the reference's streams), --out PREFIX (for the image_error files).
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>

#include "visgeom_b200/calibration.hpp"

int main(int argc, char **argv)
{
    visgeom_b200::GenericCameraCalibration calibration;
    try {
        int i = 1;
        for (; i + 1 < argc && argv[i][0] == '-'; i += 2) {
            if (!strcmp(argv[i], "--device")) calibration.device = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--precision")) std::cout.precision(atoi(argv[i + 1]));
            else if (!strcmp(argv[i], "--out")) calibration.outputPrefix = argv[i + 1];
            else { std::cerr << "unknown option " << argv[i] << std::endl; return 2; }
        }
        if (i >= argc) { std::cerr << "usage: vg_calib [--device N] [--precision N] [--out PREFIX] problem.json ..." << std::endl; return 2; }
        for (; i < argc; i++) calibration.addResiduals(argv[i]);
        calibration.compute();
    } catch (const std::exception &e) {
        // the reference lets runtime_error escape main (terminate); same text, but a clean exit code
        std::cerr << "vg_calib: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
