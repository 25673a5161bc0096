// calibration.cpp -- see include/visgeom_b200/calibration.hpp.  Host-side mirror of GenericCameraCalibration
// (src/calibration/unified_calibration.cpp) over the C ABI of the CUDA engine.
#include "visgeom_b200/calibration.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <thread>

#include "image_io.hpp"
#include "json.hpp"
#include "visgeom_b200/corner_detector.hpp"

namespace visgeom_b200 {

using std::cout;
using std::endl;
using std::runtime_error;
using std::string;
using std::vector;

namespace {

// transformFromData, include/json.h:37-68
Transf transformFromData(const vector<double> &v)
{
    if (v.size() == 3) return Transf(v[0], v[1], 0, 0, 0, v[2]);                 // x, y, theta
    if (v.size() == 6) return Transf(v.data());                                   // angle-axis
    if (v.size() == 7) return Transf(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);   // quaternion
    if (v.size() == 12) {                                                         // homogeneous 3x4
        const Matrix3d R{{v[0], v[1], v[2], v[4], v[5], v[6], v[8], v[9], v[10]}};
        return Transf(Vector3d(v[3], v[7], v[11]), R);
    }
    throw runtime_error("invalid trasformation format. must be 3, 6, or 12 values; " + std::to_string(v.size()) +
                        " are given.");
}

void check(int rc, const char *what)
{
    if (rc < 0) throw runtime_error(string(what) + ": " + vg_last_error());
}

struct ProblemHandle {
    vg_problem *p;
    explicit ProblemHandle(int device) : p(vg_problem_create(device))
    {
        if (!p) throw runtime_error(string("vg_problem_create: ") + vg_last_error());
    }
    ~ProblemHandle() { vg_problem_destroy(p); }
};

}  // namespace

GenericCameraCalibration::~GenericCameraCalibration()
{
    for (auto &x : cameraMap) delete x.second;
}

bool GenericCameraCalibration::addResiduals(const string &infoFileName)
{
    const json::Value root = json::read_json(infoFileName);
    parseTransforms(root);
    parseCameras(root);
    parseData(root);
    return true;
}

// unified_calibration.cpp:91-132
void GenericCameraCalibration::parseTransforms(const json::Value &root)
{
    for (const json::Value &node : root.child("transformations").arr) {
        const string name = node.getString("name");
        TransformInfo &info = transformInfoMap[name] = TransformInfo();
        info.global = node.getBool("global");
        info.prior = node.getBool("prior");
        info.constant = node.getBool("constant");
        if (info.constant && !info.prior) throw runtime_error(name + " is constant but there is no prior");
        if (info.global) {
            globalTransformMap[name] = Array6d{};
            if (info.prior) globalTransformMap[name] = transformFromData(node.child("value").numbers()).toArray();
        } else {
            sequenceTransformMap[name] = vector<Array6d>();
            sequenceInitMap[name] = vector<bool>();
            if (info.prior)
                for (const json::Value &val : node.child("value").arr)
                    sequenceTransformMap[name].push_back(transformFromData(val.numbers()).toArray());
        }
    }
}

// unified_calibration.cpp:134-180
void GenericCameraCalibration::parseCameras(const json::Value &root)
{
    for (const json::Value &node : root.child("cameras").arr) {
        const string name = node.getString("name");
        cameraConstantMap[name] = node.getBool("constant");
        vector<double> &intrinsicVec = intrinsicMap[name] = node.child("value").numbers();
        const string cameraType = node.getString("type");
        const size_t n = intrinsicVec.size();
        if (cameraType == "eucm") {
            cout << "Model : EUCM" << endl;
            if (n != 6) throw runtime_error("invalid number of intrinsic parameters");
            cameraMap[name] = new EnhancedCamera(intrinsicVec.data());
        } else if (cameraType == "ucm") {
            cout << "Model : UCM" << endl;
            if (n != 5) throw runtime_error("invalid number of intrinsic parameters");
            cameraMap[name] = new UnifiedCamera(intrinsicVec.data());
        } else if (cameraType == "mei") {
            cout << "Model : MEI" << endl;
            if (n != 10) throw runtime_error("invalid number of intrinsic parameters");
            cameraMap[name] = new MeiCamera(intrinsicVec.data());
        } else {
            throw runtime_error("invalid camera model name");
        }
    }
}

// unified_calibration.cpp:182-230
void GenericCameraCalibration::initTransformChainInfo(ImageData &data, const json::Value &node)
{
    data.cameraName = node.getString("camera");
    if (cameraMap.find(data.cameraName) == cameraMap.end()) throw runtime_error(data.cameraName + " : unknown camera");
    for (const json::Value &flag : node.child("parameters").arr) {
        const string &f = flag.str;
        if (f == "do_not_solve") data.doNotSolve = true;
        else if (f == "do_not_solve_global") data.doNotSolveGlobal = true;
        else if (f == "show_outliers") data.showOutliers = true;
        else if (f == "improve_detection") data.improveDetection = true;
        else if (f == "check_extraction" || f == "user_guided" ||
                 f == "save_outlire_images" || f == "draw_improved") { /* GUI options: nothing to do here */ }
        else cout << "WARNING : UNKNOWN FLAG -- " << f << endl;
    }
    cout << "Camera : " << data.cameraName << endl;
    cout << "Transformations : ";
    for (const json::Value &t : node.child("transform_chain").arr) {
        data.transNameVec.push_back(t.getString("name"));
        cout << data.transNameVec.back();
        if (t.getBool("direct")) data.transStatusVec.push_back(TRANSFORM_DIRECT);
        else { data.transStatusVec.push_back(TRANSFORM_INVERSE); cout << "_inv"; }
        cout << "   ";
    }
    int sequenceCount = 0;
    for (const string &name : data.transNameVec) {
        if (transformInfoMap.find(name) == transformInfoMap.end()) throw runtime_error(name + " has not been declared");
        if (!transformInfoMap[name].global) sequenceCount++;
    }
    if (sequenceCount != 1) throw runtime_error("not one sequences in a transform chain");
    if (data.transNameVec.size() > VG_MAX_CHAIN)
        throw runtime_error("the transform chain is too long (5 transforms at max are supproted)");
    cout << endl;
}

// unified_calibration.cpp:233-250
void GenericCameraCalibration::initGridIR(ImageData &data, const json::Value &node)
{
    data.board.clear();
    data.idxUL = node.getInt("object.corner_ul");
    data.idxUR = node.getInt("object.corner_ur");
    data.idxBL = node.getInt("object.corner_bl");
    data.idxBR = node.getInt("object.corner_br");
    for (const json::Value &x : node.child("object.points").arr) {
        const vector<double> pt = x.numbers();
        data.board.emplace_back(pt.at(0), pt.at(1), pt.at(2));
    }
}

// unified_calibration.cpp:252-277
void GenericCameraCalibration::readCorners(ImageData &data, const json::Value &node)
{
    data.imageWidth = node.getInt("image_width");
    data.imageHeight = node.getInt("image_height");
    const json::Value dataFile = json::read_json(node.getString("data_file"));
    const string cameraID = node.getString("camera");
    for (const json::Value &dataPoint : dataFile.arr) {
        data.detectedCornersVec.emplace_back();
        vector<Vector2d> &cornerVec = data.detectedCornersVec.back();
        for (const json::Value &x : dataPoint.arr) {
            if (x.getString("camera") == cameraID) {
                for (const json::Value &y : x.child("points").arr) {
                    const vector<double> pt = y.numbers();
                    cornerVec.emplace_back(pt.at(0), pt.at(1));
                }
                break;
            }
        }
        if (!cornerVec.empty() && cornerVec.size() != data.board.size())
            throw runtime_error("an image has " + std::to_string(cornerVec.size()) + " points, the object " +
                                std::to_string(data.board.size()));
    }
}

// unified_calibration.cpp:279-309
void GenericCameraCalibration::initGrid(ImageData &data, const json::Value &node)
{
    data.Nx = node.getInt("object.cols");
    data.Ny = node.getInt("object.rows");
    data.sqSize = node.getDouble("object.size");
    data.board.clear();
    for (int i = 0; i < data.Ny; i++)
        for (int j = 0; j < data.Nx; j++) data.board.emplace_back(data.sqSize * j, data.sqSize * i, 0);
    data.idxUL = 0;
    data.idxUR = data.Nx - 1;
    data.idxBL = data.Nx * (data.Ny - 1);
    data.idxBR = data.Nx * data.Ny - 1;
    data.useImages = true;
    const string prefix = node.getString("images.prefix");
    data.detectedCornersVec.clear();
    for (const json::Value &x : node.child("images.names").arr) data.imageNameVec.push_back(prefix + x.str);
    extractGridProjections(data);
}

// unified_calibration.cpp:992-1062: the reference runs CornerDetector image by image; here the images are read first and
// the ones of one size go through the detector together (CornerDetector::detectPatterns -> vg_detect_pattern)
void GenericCameraCalibration::extractGridProjections(ImageData &data)
{
    CornerDetector detector(data.Nx, data.Ny, 3, data.improveDetection);
    string sequenceName;
    for (const string &name : data.transNameVec)
        if (!transformInfoMap[name].global) { sequenceName = name; break; }
    const bool initialized = transformInfoMap[sequenceName].initialized;
    const vector<bool> &initVec = sequenceInitMap[sequenceName];
    const int n = (int)data.imageNameVec.size();
    data.detectedCornersVec.assign(n, {});
    vector<Mat8u> frames(n);
    vector<string> error(n);
    for (int i = 0; i < n; i++)
        if (initialized && !(i < (int)initVec.size() && initVec[i]))
            error[i] = " : ERROR, the pattern has not been found on the corresponding image";
    {   // decoding is the slow part of this function once the detector runs on the GPU: all cores read pictures
        std::atomic<int> next(0);
        auto reader = [&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1))
                if (error[i].empty()) frames[i] = image_io::imread_grey(data.imageNameVec[i]);
        };
        const int nt = std::max(1, std::min(n, (int)std::thread::hardware_concurrency()));
        vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(reader);
        reader();
        for (std::thread &t : pool) t.join();
    }
    for (int i = 0; i < n; i++) {
        if (!error[i].empty()) continue;
        if (frames[i].empty()) error[i] = " : ERROR, file not found";
        else if (data.imageWidth == 0) { data.imageWidth = frames[i].cols; data.imageHeight = frames[i].rows; }
    }
    // batches of equal size, in file order within a batch
    std::map<std::pair<int, int>, vector<int>> bySize;
    for (int i = 0; i < n; i++)
        if (!frames[i].empty()) bySize[{frames[i].cols, frames[i].rows}].push_back(i);
    for (const auto &group : bySize) {
        vector<const Mat8u *> batch;
        for (int i : group.second) batch.push_back(&frames[i]);
        const vector<vector<Vector2d>> grids = detector.detectPatterns(batch);
        for (size_t k = 0; k < group.second.size(); k++) {
            const int i = group.second[k];
            if (grids[k].empty()) error[i] = " : ERROR, pattern not found";
            else data.detectedCornersVec[i] = grids[k];
        }
    }
    int countSuccess = 0;
    for (int i = 0; i < n; i++) {
        cout << data.imageNameVec[i] << endl;
        if (!error[i].empty()) cout << data.imageNameVec[i] << error[i] << endl;
        else countSuccess++;
    }
    cout << endl << "DETECTION RATE : " << countSuccess << " of " << n << " detected" << endl;
}

Transf GenericCameraCalibration::getTransform(const string &name, int idx) const
{
    const auto g = globalTransformMap.find(name);
    if (g != globalTransformMap.end()) return Transf(g->second.data());
    return Transf(sequenceTransformMap.at(name).at(idx).data());
}

// unified_calibration.cpp:311-348: un-wind the chain around the transform being initialised
Transf GenericCameraCalibration::getInitTransform(Transf xi, const string &initName, const ImageData &data, int transfIdx)
{
    const int n = (int)data.transNameVec.size();
    for (int i = 0; i < n; i++) {
        const string &name = data.transNameVec[i];
        if (name == initName) break;
        if (data.transStatusVec[i] == TRANSFORM_DIRECT) xi = getTransform(name, transfIdx).inverseCompose(xi);
        else xi = getTransform(name, transfIdx).compose(xi);
    }
    for (int i = n - 1; i >= 0; i--) {
        const string &name = data.transNameVec[i];
        if (name == initName) {
            if (data.transStatusVec[i] == TRANSFORM_INVERSE) xi = xi.inverse();
            break;
        }
        if (data.transStatusVec[i] == TRANSFORM_DIRECT) xi = xi.composeInverse(getTransform(name, transfIdx));
        else xi = xi.compose(getTransform(name, transfIdx));
    }
    return xi;
}

// closed-form part of estimateInitialGrid, unified_calibration.cpp:1066-1129: the four outer corners are
// back-projected, scaled with the board's edge lengths and turned into a board frame
Transf GenericCameraCalibration::estimateInitialGridGuess(const ImageData &data, int gridIdx) const
{
    const vector<Vector2d> &cornerVec = data.detectedCornersVec[gridIdx];
    const ICamera *cam = cameraMap.at(data.cameraName);
    Vector3d XUL, XUR, XBL, XBR;
    cam->reconstructPoint(cornerVec[data.idxUL], XUL);
    cam->reconstructPoint(cornerVec[data.idxUR], XUR);
    cam->reconstructPoint(cornerVec[data.idxBL], XBL);
    cam->reconstructPoint(cornerVec[data.idxBR], XBR);
    XUL.normalize(); XUR.normalize(); XBR.normalize(); XBL.normalize();
    const double scaleXU = (data.board[data.idxUR] - data.board[data.idxUL]).norm() / (XUR - XUL).norm();
    const double scaleXB = (data.board[data.idxBR] - data.board[data.idxBL]).norm() / (XBR - XBL).norm();
    const double scaleYL = (data.board[data.idxBL] - data.board[data.idxUL]).norm() / (XBL - XUL).norm();
    const double scaleYR = (data.board[data.idxBR] - data.board[data.idxUR]).norm() / (XBR - XUR).norm();
    const Vector3d pos = XUL * std::min(scaleXU, scaleYL);
    const Vector3d posx = XUR * std::min(scaleXU, scaleYR);
    const Vector3d posy = XBL * std::min(scaleXB, scaleYL);
    Vector3d ex = posx - pos, ey = posy - pos;
    ex.normalize();
    ey = ey - ex * ex.dot(ey);      // make ey perpendicular
    ey.normalize();
    const Vector3d ez = ex.cross(ey);
    return Transf(pos, Matrix3d::fromColumns(ex, ey, ez));
}

// The refinement part of estimateInitialGrid (unified_calibration.cpp:1131-1155), batched: the reference solves
// one 6-parameter problem per image with the intrinsics constant; here all images of the dataset are the poses
// at once on the GPU but as independent problems, one per image as in the reference (own trust region, own accept /
// reject, own termination; vg_refine_poses), each block under SoftLOneLoss(25) (:1143).
void GenericCameraCalibration::refineInitialGrids(const ImageData &data, const vector<int> &idx, vector<Array6d> &xi) const
{
    if (idx.empty()) return;
    const ICamera *cam = cameraMap.at(data.cameraName);
    const int P = (int)data.board.size(), n = (int)idx.size();
    vector<double> board(3 * (size_t)P), obs(2 * (size_t)P * n), poses(6 * (size_t)n);
    for (int i = 0; i < P; i++) for (int k = 0; k < 3; k++) board[3 * i + k] = data.board[i][k];
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < P; i++) for (int k = 0; k < 2; k++)
            obs[((size_t)j * P + i) * 2 + k] = data.detectedCornersVec[idx[j]][i][k];
        for (int k = 0; k < 6; k++) poses[6 * (size_t)j + k] = xi[j][k];
    }
    vg_solve_options o;
    vg_solve_options_default(&o);
    o.max_num_iterations = 500;                // :1148
    o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;   // Ceres defaults (:1146-1150 set nothing else)
    // one independent problem per image, as the reference builds them: camera constant, SoftLOneLoss(25) (:1143-1145)
    check(vg_refine_poses(cam->model(), intrinsicMap.at(data.cameraName).data(), n, P, board.data(), obs.data(), poses.data(),
                          25.0, &o, nullptr, nullptr, nullptr), "vg_refine_poses (initial poses)");
    for (int j = 0; j < n; j++) for (int k = 0; k < 6; k++) xi[j][k] = poses[6 * (size_t)j + k];
}

// unified_calibration.cpp:358-427: a global transform initialised from one image is refined over all images
// of the dataset with every other block constant, every block under SoftLOneLoss(1) (:379-404)
void GenericCameraCalibration::initGlobalTransform(const ImageData &data, const string &name)
{
    const ICamera *cam = cameraMap.at(data.cameraName);
    const int P = (int)data.board.size();
    vector<int> seqIndex;
    vector<double> board(3 * (size_t)P), obs;
    for (int i = 0; i < P; i++) for (int k = 0; k < 3; k++) board[3 * i + k] = data.board[i][k];
    for (size_t t = 0; t < data.detectedCornersVec.size(); t++) {
        if (data.detectedCornersVec[t].empty()) continue;
        seqIndex.push_back((int)t);
        for (const Vector2d &c : data.detectedCornersVec[t]) { obs.push_back(c[0]); obs.push_back(c[1]); }
    }
    ProblemHandle h(device);
    const int camId = vg_problem_add_camera(h.p, cam->model(), intrinsicMap.at(data.cameraName).data(), 1);
    check(camId, "vg_problem_add_camera");
    vector<int> ids, status;
    int target = -1;
    for (size_t i = 0; i < data.transNameVec.size(); i++) {
        const string &tn = data.transNameVec[i];
        int id;
        if (transformInfoMap[tn].global) id = vg_problem_add_transform(h.p, 1, tn != name, 1, globalTransformMap[tn].data());
        else {
            const vector<Array6d> &seq = sequenceTransformMap[tn];
            id = vg_problem_add_transform(h.p, 0, 1, (int)seq.size(), seq[0].data());
        }
        check(id, "vg_problem_add_transform");
        if (tn == name) target = id;
        ids.push_back(id);
        status.push_back(data.transStatusVec[i]);
    }
    const int ds = vg_problem_add_dataset(h.p, camId, P, board.data(), (int)seqIndex.size(), obs.data(), seqIndex.data(),
                                          (int)ids.size(), ids.data(), status.data());
    check(ds, "vg_problem_add_dataset");
    check(vg_problem_set_loss(h.p, ds, 1.0), "vg_problem_set_loss");         // new SoftLOneLoss(1), :379-404
    vg_solve_options o;
    vg_solve_options_default(&o);
    o.max_num_iterations = 500;
    o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;
    o.verbose = 1;
    vg_solve_summary s;
    check(vg_problem_solve(h.p, &o, &s), "vg_problem_solve (global transform)");
    check(vg_problem_get_transform(h.p, target, globalTransformMap[name].data()), "vg_problem_get_transform");
}

// unified_calibration.cpp:429-512
void GenericCameraCalibration::initTransforms(const ImageData &data, const string &initName)
{
    if (initName == "none") return;
    if (transformInfoMap.find(initName) == transformInfoMap.end())
        throw runtime_error(initName + " does not exist, impossible to initialize");
    if (std::find(data.transNameVec.begin(), data.transNameVec.end(), initName) == data.transNameVec.end())
        throw runtime_error(initName + " does not belong to the transform chain");
    if (transformInfoMap[initName].prior) throw runtime_error(initName + " has a prior value");
    transformInfoMap[initName].initialized = true;
    for (const string &x : data.transNameVec)
        if (!(transformInfoMap[x].prior ^ transformInfoMap[x].initialized))
            throw runtime_error(x + " is not initialized. Cannot initialize more than one transform at a time");

    const int nImg = (int)data.detectedCornersVec.size();
    if (!transformInfoMap[initName].global) {
        vector<Array6d> &seq = sequenceTransformMap[initName];
        vector<bool> &done = sequenceInitMap[initName];
        if (seq.empty()) { seq.assign(nImg, Array6d{0, 0, 1, 0, 0, 0}); done.assign(nImg, false); }
        if ((int)seq.size() < nImg) throw runtime_error(initName + " : the sequence is shorter than the dataset");
        vector<int> idx;
        vector<Array6d> xi;
        for (int t = 0; t < nImg; t++)
            if (!data.detectedCornersVec[t].empty() && !done[t]) {
                idx.push_back(t);
                xi.push_back(estimateInitialGridGuess(data, t).toArray());
            }
        if (!data.doNotSolve) refineInitialGrids(data, idx, xi);
        for (size_t j = 0; j < idx.size(); j++) {
            seq[idx[j]] = getInitTransform(Transf(xi[j].data()), initName, data, idx[j]).toArray();
            done[idx[j]] = true;
        }
    } else {
        const int t = data.getFirstExtractedIdx();
        if (t < 0) throw runtime_error(initName + " : no board extracted, impossible to initialize");
        vector<Array6d> xi{estimateInitialGridGuess(data, t).toArray()};
        if (!data.doNotSolve) refineInitialGrids(data, {t}, xi);
        cout << "INITI VALUE IN CAMERA FRAME " << endl << Transf(xi[0].data()) << endl;
        const Transf x = getInitTransform(Transf(xi[0].data()), initName, data, t);
        cout << "INITI TRANSFORM " << endl << x << endl;
        globalTransformMap[initName] = x.toArray();
        if (nImg > 1 && !data.doNotSolve) initGlobalTransform(data, initName);
    }
}

// unified_calibration.cpp:632-660 (+ the dataset types this engine does not take)
void GenericCameraCalibration::parseData(const json::Value &root)
{
    for (const json::Value &node : root.child("data").arr) {
        const string dataType = node.getString("type");
        if (dataType == "ir_data") {
            dataVec.emplace_back();
            ImageData &data = dataVec.back();
            initTransformChainInfo(data, node);
            initGridIR(data, node);
            readCorners(data, node);
            initTransforms(data, node.getString("init"));
        } else if (dataType == "odometry") {
            parseOdometry(node);
        } else if (dataType == "transformation_prior") {
            parseTransformationPrior(node);
        } else if (dataType == "images") {
            dataVec.emplace_back();
            ImageData &data = dataVec.back();
            initTransformChainInfo(data, node);
            initGrid(data, node);
            initTransforms(data, node.getString("init"));
        } else {
            throw runtime_error("dataset type \"" + dataType + "\" is not supported by this engine");
        }
    }
}

// unified_calibration.cpp:742-807: odometry readings -> one OdometryPrior block per pair of consecutive elements
void GenericCameraCalibration::parseOdometry(const json::Value &node)
{
    OdometryData od;
    od.transformName = node.getString("transform");
    const string &name = od.transformName;
    if (transformInfoMap.find(name) == transformInfoMap.end()) throw runtime_error(name + " has not been declared");
    if (transformInfoMap[name].global) throw runtime_error(name + " is global. Odometry must be a sequence");
    od.errV = node.getDouble("err_v");        // relative error in speed
    od.errW = node.getDouble("err_w");        // relative error in rotation
    od.lambda = node.getDouble("lambda");
    for (const json::Value &item : node.child("value").arr) od.odometry.push_back(transformFromData(item.numbers()).toArray());
    if (node.getBool("init")) {               // use the odometry as initial values (:776-790)
        cout << name << endl;
        if (!sequenceTransformMap[name].empty()) throw runtime_error(name + " has already been initialized");
        transformInfoMap[name].initialized = true;
        sequenceTransformMap[name] = od.odometry;
        sequenceInitMap[name].assign(od.odometry.size(), true);
    }
    od.anchor = node.getBool("anchor");
    odometryVec.push_back(od);
}

// unified_calibration.cpp:808-829
void GenericCameraCalibration::parseTransformationPrior(const json::Value &node)
{
    PriorData pr;
    pr.transformName = node.getString("transform");
    const string &name = pr.transformName;
    if (transformInfoMap.find(name) == transformInfoMap.end()) throw runtime_error(name + " has not been declared");
    if (!transformInfoMap[name].prior) throw runtime_error(name + " must have a prior value");
    const vector<double> st = node.child("stiffness").numbers();
    if (st.size() != 6) throw runtime_error(name + " : the stiffness must have 6 values");
    for (int k = 0; k < 6; k++) pr.stiffness[k] = st[k];
    // getTransformData(name): the global value, or the first element of a sequence (unified_calibration.h:161-165)
    if (transformInfoMap[name].global) pr.prior = globalTransformMap[name];
    else {
        if (sequenceTransformMap[name].empty()) throw runtime_error(name + " has no values");
        pr.prior = sequenceTransformMap[name][0];
    }
    priorVec.push_back(pr);
}

// unified_calibration.cpp:1160-1183
void GenericCameraCalibration::computeTransforms(const ImageData &data, vector<Transf> &transfVec) const
{
    transfVec.clear();
    for (size_t t = 0; t < data.detectedCornersVec.size(); t++) {
        Transf xi(0, 0, 0, 0, 0, 0);
        for (size_t i = 0; i < data.transNameVec.size(); i++) {
            const Transf x = getTransform(data.transNameVec[i], (int)t);
            xi = data.transStatusVec[i] == TRANSFORM_DIRECT ? xi.compose(x) : xi.composeInverse(x);
        }
        transfVec.push_back(xi);
    }
}

// unified_calibration.cpp:1186-1218: "err_u err_v   proj_u proj_v   tx ty tz rx ry rz" per corner, err = detected -
// projected; both come from the engine's residuals (r = projected - detected)
void GenericCameraCalibration::writeImageResidual(vg_problem *p, int dataset, const ImageData &data, const string &fileName) const
{
    vector<Transf> transfVec;
    computeTransforms(data, transfVec);
    const int P = (int)data.board.size();
    int n = 0;
    for (const auto &c : data.detectedCornersVec) if (!c.empty()) n++;
    vector<double> r((size_t)n * 2 * P);
    if (n) check(vg_problem_residuals(p, dataset, r.data()), "vg_problem_residuals");
    std::ofstream f(fileName);
    f.precision(cout.precision());      // the reference's default (6 digits) unless --precision asked for more
    size_t row = 0;
    for (size_t t = 0; t < transfVec.size(); t++) {
        if (data.detectedCornersVec[t].empty()) continue;
        const size_t row0 = row;
        double stdAcc = 0;
        for (int i = 0; i < P; i++, row++) {
            const double ru = r[2 * row], rv = r[2 * row + 1];
            const Vector2d &det = data.detectedCornersVec[t][i];
            f << -ru << " " << -rv << "   " << det[0] + ru << " " << det[1] + rv << "   " << transfVec[t] << "\n";
            stdAcc += ru * ru + rv * rv;
        }
        if (!data.showOutliers) continue;
        // :1217-1232, 1234-1272 without the image: a corner is an outlier when its error reaches 3.6 sigma or one pixel
        const double sigma = std::sqrt(stdAcc / (P - 2));
        bool header = false;
        for (int i = 0; i < P; i++) {
            const double ru = r[2 * (row0 + i)], rv = r[2 * (row0 + i) + 1];
            const double errNorm = std::sqrt(ru * ru + rv * rv);
            if (errNorm < 3.6 * sigma && errNorm < 1.) continue;
            if (!header) {
                cout << "Sample #" << t << endl << transfVec[t] << endl << "standard deviation : " << sigma << endl;
                header = true;
            }
            const Vector2d &det = data.detectedCornersVec[t][i];
            cout << det[0] + ru << " " << det[1] + rv << "   err : " << errNorm << endl;
        }
    }
}

// unified_calibration.cpp:39-89 with ceres::Problem / ceres::Solve replaced by the engine
// (problem assembly: addGridResidualBlocks :514-630)
bool GenericCameraCalibration::compute()
{
    ProblemHandle h(device);
    std::map<string, int> camId, trId;
    for (auto &x : cameraMap) {
        const int id = vg_problem_add_camera(h.p, x.second->model(), intrinsicMap[x.first].data(), cameraConstantMap[x.first]);
        check(id, "vg_problem_add_camera");
        camId[x.first] = id;
    }
    for (auto &x : globalTransformMap) {
        const int id = vg_problem_add_transform(h.p, 1, transformInfoMap[x.first].constant, 1, x.second.data());
        check(id, "vg_problem_add_transform");
        trId[x.first] = id;
    }
    for (auto &x : sequenceTransformMap) {
        if (x.second.empty()) continue;
        const int id = vg_problem_add_transform(h.p, 0, transformInfoMap[x.first].constant, (int)x.second.size(), x.second[0].data());
        check(id, "vg_problem_add_transform");
        trId[x.first] = id;
    }
    vector<int> dsId(dataVec.size(), -1);
    for (size_t d = 0; d < dataVec.size(); d++) {
        const ImageData &data = dataVec[d];
        if (data.doNotSolveGlobal) continue;
        const int P = (int)data.board.size();
        vector<double> board(3 * (size_t)P), obs;
        vector<int> seqIndex, ids, status;
        for (int i = 0; i < P; i++) for (int k = 0; k < 3; k++) board[3 * i + k] = data.board[i][k];
        for (size_t t = 0; t < data.detectedCornersVec.size(); t++) {
            if (data.detectedCornersVec[t].empty()) continue;       // :520
            seqIndex.push_back((int)t);
            for (const Vector2d &c : data.detectedCornersVec[t]) { obs.push_back(c[0]); obs.push_back(c[1]); }
        }
        for (size_t i = 0; i < data.transNameVec.size(); i++) {
            if (trId.find(data.transNameVec[i]) == trId.end())
                throw runtime_error(data.transNameVec[i] + " has no values: give a prior or initialise it");
            ids.push_back(trId[data.transNameVec[i]]);
            status.push_back(data.transStatusVec[i]);
        }
        dsId[d] = vg_problem_add_dataset(h.p, camId[data.cameraName], P, board.data(), (int)seqIndex.size(), obs.data(),
                                         seqIndex.data(), (int)ids.size(), ids.data(), status.data());
        check(dsId[d], "vg_problem_add_dataset");
    }

    for (const PriorData &pr : priorVec) {
        if (trId.find(pr.transformName) == trId.end()) throw runtime_error(pr.transformName + " has no values");
        check(vg_problem_add_transformation_prior(h.p, trId[pr.transformName], 0, pr.stiffness.data(), pr.prior.data()),
              "vg_problem_add_transformation_prior");
    }
    for (const OdometryData &od : odometryVec) {
        if (trId.find(od.transformName) == trId.end())
            throw runtime_error(od.transformName + " has no values: give a prior or initialise it");
        if (od.odometry.size() != sequenceTransformMap[od.transformName].size())
            throw runtime_error(od.transformName + " : " + std::to_string(od.odometry.size()) + " odometry readings for a sequence of " +
                                std::to_string(sequenceTransformMap[od.transformName].size()) + " elements");
        check(vg_problem_add_odometry(h.p, trId[od.transformName], od.errV, od.errW, od.lambda, (int)od.odometry.size(),
                                      od.odometry[0].data()), "vg_problem_add_odometry");
        if (od.anchor) check(vg_problem_set_pose_constant(h.p, trId[od.transformName], 0, 1), "vg_problem_set_pose_constant");
    }

    vg_solve_options o;
    vg_solve_options_default(&o);      // max_num_iterations 1000, tolerances 1e-15 (:46-49)
    o.verbose = 1;                     // minimizer_progress_to_stdout
    check(vg_problem_solve(h.p, &o, &lastSummary), "vg_problem_solve");
    static const char *why[] = {"function tolerance", "gradient tolerance", "parameter tolerance", "max iterations",
                                "trust region radius", "failure"};
    cout << "\nSolver Summary\n  iterations " << lastSummary.iterations << " (successful " << lastSummary.num_successful
         << ", unsuccessful " << lastSummary.num_unsuccessful << ")\n  initial cost " << lastSummary.initial_cost
         << "\n  final cost   " << lastSummary.final_cost << "\n  termination  " << why[lastSummary.termination]
         << "\n  time total " << lastSummary.seconds_total << " s, in residual/Jacobian/normal-equation kernels "
         << lastSummary.seconds_evaluate << " s (" << lastSummary.num_evaluations << " evaluations)\n" << endl;

    // results back into the maps
    for (auto &x : camId) check(vg_problem_get_camera(h.p, x.second, intrinsicMap[x.first].data()), "vg_problem_get_camera");
    for (auto &x : globalTransformMap) check(vg_problem_get_transform(h.p, trId[x.first], x.second.data()), "vg_problem_get_transform");
    for (auto &x : sequenceTransformMap)
        if (!x.second.empty()) check(vg_problem_get_transform(h.p, trId[x.first], x.second[0].data()), "vg_problem_get_transform");
    for (auto &x : cameraMap) x.second->setParameters(intrinsicMap[x.first].data());

    cout << "Intrinsic parameters :" << endl;
    for (auto &x : intrinsicMap) {
        cout << x.first << " : ";
        for (double v : x.second) cout << v << "  ";
        cout << endl;
    }
    cout << "Local extrinsic parameters :" << endl;
    for (auto &seq : sequenceTransformMap) {
        cout << "Sequence : " << seq.first << endl;
        int i = 0;
        for (auto &x : seq.second) cout << i++ << " : " << Transf(x.data()) << endl;
    }
    cout << "Global extrinsic parameters :" << endl;
    for (auto &x : globalTransformMap) cout << x.first << " : " << Transf(x.second.data()) << endl;

    for (size_t d = 0; d < dataVec.size(); d++)
        if (dsId[d] >= 0) writeImageResidual(h.p, dsId[d], dataVec[d], outputPrefix + "image_error_" + std::to_string(d) + ".txt");
    return true;
}

}  // namespace visgeom_b200
