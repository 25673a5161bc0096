// calibration.hpp -- GenericCameraCalibration with visgeom's interface
// (include/calibration/unified_calibration.h, src/calibration/unified_calibration.cpp), running on the CUDA engine:
// addResiduals() reads the reference's JSON problem files, compute() replaces ceres::Solve + the report.
//
// Supported dataset types: "ir_data" (pre-extracted corners, unified_calibration.cpp:234-277,648-660), "odometry"
// (OdometryPrior blocks between consecutive elements of a sequence, :742-807) and "transformation_prior" (:808-829).
// The "images" type needs the checkerboard detector (OpenCV), which is outside this engine, and is rejected with a
// message; so is "odometry_intrinsic" (wheel-odometry intrinsics as parameters, :661-741).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../visgeom_b200.h"
#include "camera.hpp"
#include "geometry.hpp"

namespace visgeom_b200 {

namespace json { struct Value; }

enum TransformationStatus { TRANSFORM_DIRECT = VG_TRANSFORM_DIRECT, TRANSFORM_INVERSE = VG_TRANSFORM_INVERSE };

struct TransformInfo {
    bool global = false, prior = false, constant = false, initialized = false;
};

struct ImageData {
    std::string cameraName;
    std::vector<std::string> transNameVec;
    std::vector<TransformationStatus> transStatusVec;
    std::vector<Vector3d> board;
    std::vector<std::vector<Vector2d>> detectedCornersVec;     // per image; empty: no board extracted
    int idxUL = 0, idxUR = 0, idxBL = 0, idxBR = 0;
    int imageWidth = 0, imageHeight = 0;
    // "images" datasets (unified_calibration.cpp:279-309): the board and where its pictures are
    int Nx = 0, Ny = 0;
    double sqSize = 0;
    bool useImages = false, improveDetection = false;
    std::vector<std::string> imageNameVec;
    bool doNotSolve = false, doNotSolveGlobal = false;
    bool showOutliers = false;      // "show_outliers": the textual part of the reference's outlier report (no image windows here)
    int getFirstExtractedIdx() const
    {
        for (size_t i = 0; i < detectedCornersVec.size(); i++) if (!detectedCornersVec[i].empty()) return (int)i;
        return -1;
    }
};

// "odometry" dataset (unified_calibration.cpp:742-807)
struct OdometryData {
    std::string transformName;
    double errV = 0, errW = 0, lambda = 0;
    std::vector<Array6d> odometry;         // one reading per element of the sequence
    bool anchor = false;                   // first element constant (:803-806)
};

// "transformation_prior" dataset (:808-829): the prior value is the transform's value when the block is created
struct PriorData {
    std::string transformName;
    Array6d stiffness{}, prior{};
};

class GenericCameraCalibration {
public:
    GenericCameraCalibration() {}
    ~GenericCameraCalibration();
    GenericCameraCalibration(const GenericCameraCalibration &) = delete;
    GenericCameraCalibration &operator=(const GenericCameraCalibration &) = delete;

    // read one problem file and add its cameras, transformations and datasets (unified_calibration.cpp:350-356)
    bool addResiduals(const std::string &infoFileName);
    // solve, print the report, write image_error_<i>.txt (unified_calibration.cpp:39-89)
    bool compute();

    // results
    const std::map<std::string, std::vector<double>> &intrinsics() const { return intrinsicMap; }
    const std::map<std::string, Array6d> &globalTransforms() const { return globalTransformMap; }
    const std::map<std::string, std::vector<Array6d>> &sequenceTransforms() const { return sequenceTransformMap; }
    const vg_solve_summary &summary() const { return lastSummary; }
    int device = -1;                 // CUDA device (-1: current)
    std::string outputPrefix;        // prepended to image_error_<i>.txt

private:
    void parseTransforms(const json::Value &root);
    void parseCameras(const json::Value &root);
    void parseData(const json::Value &root);
    void parseOdometry(const json::Value &node);
    void parseTransformationPrior(const json::Value &node);
    void initTransformChainInfo(ImageData &data, const json::Value &node);
    void initGridIR(ImageData &data, const json::Value &node);
    void readCorners(ImageData &data, const json::Value &node);
    void initGrid(ImageData &data, const json::Value &node);
    void extractGridProjections(ImageData &data);
    void initTransforms(const ImageData &data, const std::string &initName);
    void initGlobalTransform(const ImageData &data, const std::string &name);
    Transf estimateInitialGridGuess(const ImageData &data, int gridIdx) const;
    void refineInitialGrids(const ImageData &data, const std::vector<int> &idx, std::vector<Array6d> &xi) const;
    Transf getInitTransform(Transf xi, const std::string &initName, const ImageData &data, int transfIdx);
    Transf getTransform(const std::string &name, int idx) const;
    void computeTransforms(const ImageData &data, std::vector<Transf> &transfVec) const;
    void writeImageResidual(vg_problem *p, int dataset, const ImageData &data, const std::string &fileName) const;

    std::map<std::string, TransformInfo> transformInfoMap;
    std::map<std::string, Array6d> globalTransformMap;
    std::map<std::string, std::vector<Array6d>> sequenceTransformMap;
    std::map<std::string, std::vector<bool>> sequenceInitMap;
    std::map<std::string, ICamera *> cameraMap;
    std::map<std::string, std::vector<double>> intrinsicMap;
    std::map<std::string, bool> cameraConstantMap;
    std::vector<ImageData> dataVec;
    std::vector<OdometryData> odometryVec;
    std::vector<PriorData> priorVec;
    vg_solve_summary lastSummary{};
};

}  // namespace visgeom_b200
