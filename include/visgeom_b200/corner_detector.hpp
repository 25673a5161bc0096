// corner_detector.hpp -- host-side mirror of visgeom's CornerDetector (include/calibration/corner_detector.h:49-60,
// src/calibration/corner_detector.cpp:200-260) over the C ABI: same constructor arguments, setImage + detectPattern
// per image as GenericCameraCalibration::extractGridProjections uses them (unified_calibration.cpp:995,1031-1033),
// plus the batched call that is the fast path on the GPU (all images of a dataset in one vg_detect_pattern).
// Images are 8-bit single-channel, row-major (cv::Mat_<uint8_t> in the reference: Mat8u).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../visgeom_b200.h"
#include "geometry.hpp"

namespace visgeom_b200 {

struct Mat8u {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    bool empty() const { return rows == 0 || cols == 0; }
    uint8_t operator()(int v, int u) const { return data[(size_t)v * cols + u]; }
};

class CornerDetector {
public:
    // initRadius and debug are accepted for source compatibility: detectPattern derives the radius from its scales
    // (corner_detector.cpp:231) and the debug windows need a display
    CornerDetector(int Nx, int Ny, int initRadius = 5, bool improveDetection = true, bool debug = false)
        : _Nx(Nx), _Ny(Ny), IMPROVE_DETECTION(improveDetection)
    {
        (void)initRadius; (void)debug;
    }

    void setImage(const Mat8u &img) { _img = img; }

    // corner_detector.cpp:223-260
    bool detectPattern(std::vector<Vector2d> &ptVec)
    {
        if (_img.empty()) return false;
        std::vector<double> c((size_t)_Nx * _Ny * 2);
        unsigned char found = 0;
        if (vg_detect_pattern(_img.data.data(), 1, _img.cols, _img.rows, _Nx, _Ny, IMPROVE_DETECTION ? 1 : 0, c.data(), &found))
            throw std::runtime_error(std::string("vg_detect_pattern: ") + vg_last_error());
        if (!found) return false;
        ptVec.clear();
        for (int i = 0; i < _Nx * _Ny; i++) ptVec.emplace_back(c[2 * i], c[2 * i + 1]);
        return true;
    }

    // all images of one size at once: result[i] is empty where no pattern was found
    std::vector<std::vector<Vector2d>> detectPatterns(const std::vector<const Mat8u *> &imgs) const
    {
        std::vector<std::vector<Vector2d>> res(imgs.size());
        if (imgs.empty()) return res;
        const int W = imgs[0]->cols, H = imgs[0]->rows, P = _Nx * _Ny;
        std::vector<uint8_t> pix((size_t)W * H * imgs.size());
        for (size_t i = 0; i < imgs.size(); i++) {
            if (imgs[i]->cols != W || imgs[i]->rows != H) throw std::runtime_error("detectPatterns: images of one batch must have one size");
            std::copy(imgs[i]->data.begin(), imgs[i]->data.end(), pix.begin() + (size_t)W * H * i);
        }
        std::vector<double> c((size_t)P * 2 * imgs.size());
        std::vector<unsigned char> found(imgs.size());
        if (vg_detect_pattern(pix.data(), (int)imgs.size(), W, H, _Nx, _Ny, IMPROVE_DETECTION ? 1 : 0, c.data(), found.data()))
            throw std::runtime_error(std::string("vg_detect_pattern: ") + vg_last_error());
        for (size_t i = 0; i < imgs.size(); i++)
            if (found[i])
                for (int k = 0; k < P; k++) res[i].emplace_back(c[((size_t)i * P + k) * 2], c[((size_t)i * P + k) * 2 + 1]);
        return res;
    }

private:
    const int _Nx, _Ny;
    const bool IMPROVE_DETECTION;
    Mat8u _img;
};

}  // namespace visgeom_b200
