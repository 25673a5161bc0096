// Test helper (CPU): runs the parsing / initialisation half of the calibration front end (addResiduals) on problem
// files whose datasets carry "do_not_solve" -- the closed-form initial poses of estimateInitialGrid and the chain
// un-winding of getInitTransform run on the host alone -- and prints the transforms it arrived at.
#include <cstdio>
#include <iostream>

#include "visgeom_b200/calibration.hpp"

int main(int argc, char **argv)
{
    try {
        visgeom_b200::GenericCameraCalibration calib;
        for (int i = 1; i < argc; i++) calib.addResiduals(argv[i]);
        for (const auto &s : calib.sequenceTransforms()) {
            printf("SEQ %s %zu\n", s.first.c_str(), s.second.size());
            for (const auto &x : s.second) printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", x[0], x[1], x[2], x[3], x[4], x[5]);
        }
        for (const auto &g : calib.globalTransforms()) {
            const auto &x = g.second;
            printf("GLOBAL %s\n%.17g %.17g %.17g %.17g %.17g %.17g\n", g.first.c_str(), x[0], x[1], x[2], x[3], x[4], x[5]);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
