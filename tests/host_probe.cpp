// Test helper (CPU): drives the host-side mirror classes of include/visgeom_b200/{geometry,camera}.hpp from
// commands on stdin, so that tests/test_host_classes.py can hold them against the reference build's golden vectors.
//   compose K a0..a5 b0..b5      -> 6 numbers   (K = 0 compose, 1 composeInverse, 2 inverseCompose)
//   rotmat r0 r1 r2              -> 9 numbers (row-major) then the rotation vector recovered from that matrix
//   quat t0 t1 t2 qx qy qz qw    -> 6 numbers: Transformation(x, y, z, qx, qy, qz, qw).toArray()
//   reconstruct M n p0..p(n-1) u v -> ok X Y Z
//   bounds M n p0..p(n-1) idx    -> lower upper
//   project M n p0..p(n-1) m X0 Y0 Z0 ..  (GPU) -> per point: ok u v dudx(3) dvdx(3) dudalpha(n) dvdalpha(n) through the
//                                  single-point virtuals, then "cloud" all_ok and per point: mask u v via projectPointCloud,
//                                  then per point: mask X Y Z via reconstructPointCloud of those image points
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "visgeom_b200/camera.hpp"

using namespace visgeom_b200;

static ICamera *make(int model, const double *p)
{
    if (model == VG_MODEL_EUCM) return new EnhancedCamera(p);
    if (model == VG_MODEL_UCM) return new UnifiedCamera(p);
    return new MeiCamera(p);
}

int main()
{
    std::string cmd;
    while (std::cin >> cmd) {
        if (cmd == "compose") {
            int k; double a[6], b[6];
            std::cin >> k;
            for (double &x : a) std::cin >> x;
            for (double &x : b) std::cin >> x;
            const Transf A(a), B(b);
            const Transf C = k == 0 ? A.compose(B) : (k == 1 ? A.composeInverse(B) : A.inverseCompose(B));
            const Array6d o = C.toArray();
            printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", o[0], o[1], o[2], o[3], o[4], o[5]);
        } else if (cmd == "rotmat") {
            double r[3];
            for (double &x : r) std::cin >> x;
            const Matrix3d R = rotationMatrix(Vector3d(r[0], r[1], r[2]));
            for (int i = 0; i < 9; i++) printf("%.17g ", R.m[i]);
            const Vector3d back = rotationVector(R);
            printf("%.17g %.17g %.17g\n", back[0], back[1], back[2]);
        } else if (cmd == "quat") {
            double v[7];
            for (double &x : v) std::cin >> x;
            const Array6d o = Transf(v[0], v[1], v[2], v[3], v[4], v[5], v[6]).toArray();
            printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", o[0], o[1], o[2], o[3], o[4], o[5]);
        } else if (cmd == "reconstruct" || cmd == "bounds") {
            int model, n;
            std::cin >> model >> n;
            std::vector<double> p(n);
            for (double &x : p) std::cin >> x;
            std::unique_ptr<ICamera> cam(make(model, p.data()));
            std::unique_ptr<ICamera> copy(cam->clone());
            if (cmd == "bounds") {
                int idx; std::cin >> idx;
                printf("%.17g %.17g\n", copy->lowerBound(idx), copy->upperBound(idx));
            } else {
                double u, v; std::cin >> u >> v;
                Vector3d X;
                const bool ok = copy->reconstructPoint(Vector2d(u, v), X);
                printf("%d %.17g %.17g %.17g\n", ok ? 1 : 0, X[0], X[1], X[2]);
            }
        } else if (cmd == "project") {
            int model, n, m;
            std::cin >> model >> n;
            std::vector<double> p(n);
            for (double &x : p) std::cin >> x;
            std::cin >> m;
            Vector3dVec X(m);
            for (auto &x : X) std::cin >> x[0] >> x[1] >> x[2];
            std::unique_ptr<ICamera> cam(make(model, p.data()));
            for (int i = 0; i < m; i++) {
                Vector2d uv(-7, -7);
                std::vector<double> ju(3), jv(3), au(n), av(n);
                const bool ok = cam->projectPoint(X[i], uv);
                const bool okj = cam->projectionJacobian(X[i], ju.data(), jv.data());
                const bool oka = cam->intrinsicJacobian(X[i], au.data(), av.data());
                printf("%d %.17g %.17g", (ok ? 1 : 0) | (okj ? 2 : 0) | (oka ? 4 : 0), uv[0], uv[1]);
                for (double x : ju) printf(" %.17g", x);
                for (double x : jv) printf(" %.17g", x);
                for (double x : au) printf(" %.17g", x);
                for (double x : av) printf(" %.17g", x);
                printf("\n");
            }
            Vector2dVec uv;
            std::vector<bool> mask;
            const bool all = cam->projectPointCloud(X, uv, mask);
            printf("%d\n", all ? 1 : 0);
            for (int i = 0; i < m; i++) printf("%d %.17g %.17g\n", mask[i] ? 1 : 0, uv[i][0], uv[i][1]);
            Vector3dVec back;
            std::vector<bool> rmask;
            cam->reconstructPointCloud(uv, back, rmask);
            for (int i = 0; i < m; i++) printf("%d %.17g %.17g %.17g\n", rmask[i] ? 1 : 0, back[i][0], back[i][1], back[i][2]);
        } else {
            fprintf(stderr, "unknown command %s\n", cmd.c_str());
            return 2;
        }
    }
    return 0;
}
