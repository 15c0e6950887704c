// TEST INFRASTRUCTURE ONLY — CPU oracle for the jueying_lio hot path.
// PARITY UNPINNED (see oracle.h).  Never linked into the product library.
//
// Restates, function by function (paths relative to /root/reference/src/jueying_lio):
//   IVox::AddPoints / Pos2Grid / GetClosestPoint      include/ivox3d/ivox3d.h:132-204,211-235,256-286
//   IVoxNode::KNNPointByCondition / distance2         include/ivox3d/ivox3d_node.hpp:12-16,140-205
//   hash_vec<3>                                       include/ivox3d/eigen_types.h:74-76
//   common::esti_plane                                include/common_lib.h:186-243
//   LaserMapping::ObsModel                            src/laser_mapping.cc:592-701
//   LaserMapping::MapIncremental / PointBodyToWorld   src/laser_mapping.cc:525-583,855-864
//   esekf::update_iterated_dyn_share_modified         include/IKFoM_toolkit/esekfom/esekfom.hpp:1526-1834
//   MTK SO3 / S2 / vect boxplus, boxminus, A_matrix   include/IKFoM_toolkit/mtk/types/{SOn,S2}.hpp, mtk/src/mtkmath.hpp
// Build: g++ -O3 -fopenmp -ffp-contract=off (no -march): SSE2, no FMA, like the
// reference (CMakeLists.txt:10-11).  OpenMP `parallel for` stands in for TBB par_unseq.
#include "oracle.h"
#include "smallmat.h"

#include <omp.h>
#include <chrono>
#include <cstdio>
#include <list>
#include <unordered_map>
#include <vector>

namespace orc {

struct MapPt { float x, y, z; int32_t ord; };
struct Key3 {
    int x, y, z;
    bool operator==(const Key3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Key3Hash {  // eigen_types.h:74-76 (bucket choice only; no effect on results)
    size_t operator()(const Key3& v) const {
        return size_t(((v.x) * 73856093) ^ ((v.y) * 471943) ^ ((v.z) * 83492791)) % 10000000;
    }
};

static const int kStencil26[27][3] = {
    {0, 0, 0},   {-1, 0, 0}, {1, 0, 0},   {0, 1, 0},   {0, -1, 0},  {0, 0, -1},  {0, 0, 1},  {1, 1, 0},  {-1, 1, 0},
    {1, -1, 0},  {-1, -1, 0}, {1, 0, 1},  {-1, 0, 1},  {1, 0, -1},  {-1, 0, -1}, {0, 1, 1},  {0, -1, 1}, {0, 1, -1},
    {0, -1, -1}, {1, 1, 1},  {-1, 1, 1},  {1, -1, 1},  {1, 1, -1},  {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};

struct Cand { double dist; int32_t rank; const MapPt* p; };
static inline bool cand_less(const Cand& a, const Cand& b) {
    return a.dist < b.dist || (a.dist == b.dist && a.rank < b.rank);
}

struct IVoxOracle {
    float res = 0.2f, inv_res = 5.0f;
    int nstencil = 7;
    size_t capacity = 1000000;
    using Node = std::pair<Key3, std::vector<MapPt>>;
    std::list<Node> cache;
    std::unordered_map<Key3, std::list<Node>::iterator, Key3Hash> grid;
    int64_t next_ord = 0;

    void configure(float resolution, int nearby, size_t cap) {
        res = resolution;
        inv_res = (float)(1.0 / (double)resolution);  // ivox3d.h:65 (double 1.0/float -> float member)
        nstencil = nearby == 0 ? 1 : nearby == 6 ? 7 : nearby == 26 ? 27 : 19;  // laser_mapping.cc:138-149
        capacity = cap;
    }
    Key3 pos2grid(float x, float y, float z) const {  // ivox3d.h:284-286
        return Key3{(int)std::round(x * inv_res), (int)std::round(y * inv_res), (int)std::round(z * inv_res)};
    }
    void add_points(const float* xyz, int64_t n, int64_t stride) {  // ivox3d.h:256-281
        for (int64_t i = 0; i < n; ++i) {
            const float* p = (const float*)((const char*)xyz + i * stride);
            MapPt mp{p[0], p[1], p[2], (int32_t)next_ord++};
            Key3 key = pos2grid(mp.x, mp.y, mp.z);
            auto it = grid.find(key);
            if (it == grid.end()) {
                cache.push_front({key, {}});
                grid.insert({key, cache.begin()});
                cache.front().second.push_back(mp);
                if (grid.size() >= capacity) {
                    grid.erase(cache.back().first);
                    cache.pop_back();
                }
            } else {
                it->second->second.push_back(mp);
                cache.splice(cache.begin(), cache, it->second);
                grid[key] = cache.begin();
            }
        }
    }
    // ivox3d.h:132-204 with the tie-break contract of SURVEY.md §7: stable
    // selection over the candidate vector, i.e. total order (dist, enumeration rank).
    int knn(float qx, float qy, float qz, int K, double max_range, Cand* out, int64_t* n_cell_pts = nullptr) const {
        std::vector<Cand> cands;
        cands.reserve(K * nstencil);
        Key3 key = pos2grid(qx, qy, qz);
        int32_t rank = 0;
        for (int s = 0; s < nstencil; ++s) {
            Key3 dk{key.x + kStencil26[s][0], key.y + kStencil26[s][1], key.z + kStencil26[s][2]};
            auto it = grid.find(dk);
            if (it == grid.end()) continue;
            const std::vector<MapPt>& pts = it->second->second;
            if (n_cell_pts) *n_cell_pts += (int64_t)pts.size();
            size_t old_size = cands.size();
            for (const MapPt& pt : pts) {  // ivox3d_node.hpp:159-169
                float dx = pt.x - qx, dy = pt.y - qy, dz = pt.z - qz;
                float d2f = (dx * dx + dy * dy) + dz * dz;
                double d = (double)d2f;
                if (d < max_range * max_range) cands.push_back(Cand{d, rank, &pt});
                ++rank;
            }
            if (old_size + K < cands.size()) {  // ivox3d_node.hpp:179-183
                std::sort(cands.begin() + old_size, cands.end(), cand_less);
                cands.resize(old_size + K);
            }
        }
        if (cands.empty()) return 0;
        std::sort(cands.begin(), cands.end(), cand_less);  // ivox3d.h:173-178
        int m = (int)std::min<size_t>(cands.size(), (size_t)K);
        for (int i = 0; i < m; ++i) out[i] = cands[i];
        return m;
    }
};

// ------------------------------------------------------------------ esti_plane
// common_lib.h:186-243.  pts: n x (x,y,z).  4-wide float dots use the SSE
// horizontal-add order (a0b0+a2b2)+(a1b1+a3b3) (SURVEY.md §8a notes).
static inline float dot4_sse(const float a[4], const float b[4]) {
    return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]);
}
static bool esti_plane(float plane[4], const float (*pts)[3], int n, float threshold) {
    if (n < 3) return false;
    float normvec[3];
    if (n == 5) {
        float A[15], b[5];
        for (int j = 0; j < 5; ++j) {
            A[j * 3 + 0] = pts[j][0]; A[j * 3 + 1] = pts[j][1]; A[j * 3 + 2] = pts[j][2];
            b[j] = -1.0f;
        }
        colpiv_qr_solve3<float>(A, 5, b, normvec);
    } else {
        double A[15], b[5], x[3];
        for (int j = 0; j < n; ++j) {
            A[j * 3 + 0] = pts[j][0]; A[j * 3 + 1] = pts[j][1]; A[j * 3 + 2] = pts[j][2];
            b[j] = -1.0;
        }
        colpiv_qr_solve3<double>(A, n, b, x);
        normvec[0] = (float)x[0]; normvec[1] = (float)x[1]; normvec[2] = (float)x[2];
    }
    float nn = std::sqrt((normvec[0] * normvec[0] + normvec[1] * normvec[1]) + normvec[2] * normvec[2]);
    plane[0] = normvec[0] / nn;
    plane[1] = normvec[1] / nn;
    plane[2] = normvec[2] / nn;
    plane[3] = (float)(1.0 / (double)nn);  // `1.0 / n` is double arithmetic narrowed on store
    for (int j = 0; j < n; ++j) {
        float temp[4] = {pts[j][0], pts[j][1], pts[j][2], 1.0f};
        if (std::fabs(dot4_sse(plane, temp)) > threshold) return false;
    }
    return true;
}

// ------------------------------------------------------------------ manifold math (double)
struct Quat { double x, y, z, w; };
static inline Quat qmul(const Quat& a, const Quat& b) {
    return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
static inline Quat qconj(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
static inline void cross3(const double a[3], const double b[3], double r[3]) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void qrot(const Quat& q, const double v[3], double r[3]) {  // Eigen _transformVector
    double qv[3] = {q.x, q.y, q.z}, uv[3], c2[3];
    cross3(qv, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(qv, uv, c2);
    for (int i = 0; i < 3; ++i) r[i] = v[i] + q.w * uv[i] + c2[i];
}
static inline void qtoR(const Quat& q, double R[9]) {  // Eigen toRotationMatrix
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
static inline void hat3(const double v[3], double M[9]) {
    M[0] = 0; M[1] = -v[2]; M[2] = v[1];
    M[3] = v[2]; M[4] = 0; M[5] = -v[0];
    M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}
static inline void mm3(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3 + 0] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
static const double kTol = 1e-11;  // MTK::tolerance<double>, mtkmath.hpp:129-131

static void cos_sinc_sqrt(double x2, double& c, double& sinc) {  // mtkmath.hpp:149-180
    static const double taylor_0 = std::numeric_limits<double>::epsilon();
    static const double taylor_2 = std::sqrt(taylor_0);
    static const double taylor_n = std::sqrt(taylor_2);
    if (x2 >= taylor_n) {
        double x = std::sqrt(x2);
        c = std::cos(x);
        sinc = std::sin(x) / x;
        return;
    }
    static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
    double cosi = 1., s = 1;
    double term = -1 / 2. * x2;
    for (int i = 0; i < 3; ++i) {
        cosi += term;
        term *= inv[2 * i];
        s += term;
        term *= -inv[2 * i + 1] * x2;
    }
    c = cosi;
    sinc = s;
}
// MTK::exp<scalar,3>(result, vec, scale) -> w ; mtkmath.hpp:248-255
static inline Quat so3_exp(const double v[3], double scale_half) {
    double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, sinc;
    cos_sinc_sqrt(scale_half * scale_half * n2, c, sinc);
    double mult = sinc * scale_half;
    return Quat{mult * v[0], mult * v[1], mult * v[2], c};
}
// SO3::log -> MTK::log(res, w, vec, 2, true) ; mtkmath.hpp:266-285
static inline void so3_log(const Quat& q, double r[3]) {
    double nv = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (nv < kTol) nv = kTol;
    double s = 2.0 / nv * std::atan(nv / q.w);
    r[0] = s * q.x; r[1] = s * q.y; r[2] = s * q.z;
}
static void A_matrix(const double v[3], double res[9]) {  // mtkmath.hpp:234-246
    double sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double n = std::sqrt(sq);
    for (int i = 0; i < 9; ++i) res[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (n < kTol) return;
    double H[9], HH[9];
    hat3(v, H);
    mm3(H, H, HH);
    double a = (1 - std::cos(n)) / sq, b = (1 - std::sin(n) / n) / sq;
    for (int i = 0; i < 9; ++i) res[i] = res[i] + a * H[i] + b * HH[i];
}

static const double kGravLen = 98090.0 / 10000.0;  // S2<double,98090,10000,1>, use-ikfom.hpp:10

struct State {  // use-ikfom.hpp:14-15
    double pos[3];
    Quat rot, offR;
    double offT[3], vel[3], bg[3], ba[3], grav[3];
};
static State load_state(const double* x) {
    State s;
    std::memcpy(s.pos, x, 24);
    s.rot = Quat{x[3], x[4], x[5], x[6]};
    s.offR = Quat{x[7], x[8], x[9], x[10]};
    std::memcpy(s.offT, x + 11, 24);
    std::memcpy(s.vel, x + 14, 24);
    std::memcpy(s.bg, x + 17, 24);
    std::memcpy(s.ba, x + 20, 24);
    std::memcpy(s.grav, x + 23, 24);
    return s;
}
static void store_state(const State& s, double* x) {
    std::memcpy(x, s.pos, 24);
    x[3] = s.rot.x; x[4] = s.rot.y; x[5] = s.rot.z; x[6] = s.rot.w;
    x[7] = s.offR.x; x[8] = s.offR.y; x[9] = s.offR.z; x[10] = s.offR.w;
    std::memcpy(x + 11, s.offT, 24);
    std::memcpy(x + 14, s.vel, 24);
    std::memcpy(x + 17, s.bg, 24);
    std::memcpy(x + 20, s.ba, 24);
    std::memcpy(x + 23, s.grav, 24);
}

// S2 (S2_typ == 1) ; S2.hpp:166-200
static void S2_Bx(const double vec[3], double Bx[6] /*3x2 row-major*/) {
    const double len = kGravLen;
    if (vec[0] + len > kTol) {
        Bx[0] = -vec[1];
        Bx[1] = -vec[2];
        Bx[2] = len - vec[1] * vec[1] / (len + vec[0]);
        Bx[3] = -vec[2] * vec[1] / (len + vec[0]);
        Bx[4] = -vec[2] * vec[1] / (len + vec[0]);
        Bx[5] = len - vec[2] * vec[2] / (len + vec[0]);
        for (int i = 0; i < 6; ++i) Bx[i] /= len;
    } else {
        for (int i = 0; i < 6; ++i) Bx[i] = 0;
        Bx[1 * 2 + 1] = -1;
        Bx[2 * 2 + 0] = 1;
    }
}
static void S2_boxplus(double vec[3], const double delta[2]) {  // S2.hpp:131-138
    double Bx[6];
    S2_Bx(vec, Bx);
    double Bu[3];
    for (int i = 0; i < 3; ++i) Bu[i] = Bx[i * 2] * delta[0] + Bx[i * 2 + 1] * delta[1];
    Quat q = so3_exp(Bu, 0.5);
    double R[9], r[3];
    qtoR(q, R);
    for (int i = 0; i < 3; ++i) r[i] = R[i * 3] * vec[0] + R[i * 3 + 1] * vec[1] + R[i * 3 + 2] * vec[2];
    vec[0] = r[0]; vec[1] = r[1]; vec[2] = r[2];
}
static void S2_boxminus(const double vec[3], const double other[3], double res[2]) {  // S2.hpp:140-158
    double c[3];
    cross3(vec, other, c);  // hat(vec)*other
    double v_sin = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    double v_cos = vec[0] * other[0] + vec[1] * other[1] + vec[2] * other[2];
    double theta = std::atan2(v_sin, v_cos);
    if (v_sin < kTol) {
        if (std::fabs(theta) > kTol) { res[0] = 3.1415926; res[1] = 0; }
        else { res[0] = 0; res[1] = 0; }
    } else {
        double Bx[6];
        S2_Bx(other, Bx);
        double hv[3];
        cross3(other, vec, hv);  // hat(other.vec)*vec
        double f = theta / v_sin;
        // res = theta / v_sin * Bx^T * hat(other)*vec   (left-to-right: scalar*Bx^T first)
        for (int j = 0; j < 2; ++j)
            res[j] = (f * Bx[0 * 2 + j]) * hv[0] + (f * Bx[1 * 2 + j]) * hv[1] + (f * Bx[2 * 2 + j]) * hv[2];
    }
}
static void S2_Nx_yy(const double vec[3], double Nx[6] /*2x3*/) {  // S2.hpp:219-223
    double Bx[6], H[9];
    S2_Bx(vec, Bx);
    hat3(vec, H);
    double f = 1 / kGravLen / kGravLen;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j)
            Nx[i * 3 + j] = (f * Bx[0 * 2 + i]) * H[0 * 3 + j] + (f * Bx[1 * 2 + i]) * H[1 * 3 + j] + (f * Bx[2 * 2 + i]) * H[2 * 3 + j];
}
static void S2_Mx(const double vec[3], const double delta[2], double Mx[6] /*3x2*/) {  // S2.hpp:225-236
    double Bx[6], H[9];
    S2_Bx(vec, Bx);
    hat3(vec, H);
    if (std::sqrt(delta[0] * delta[0] + delta[1] * delta[1]) < kTol) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j)
                Mx[i * 2 + j] = -H[i * 3 + 0] * Bx[0 * 2 + j] + -H[i * 3 + 1] * Bx[1 * 2 + j] + -H[i * 3 + 2] * Bx[2 * 2 + j];
    } else {
        double Bu[3];
        for (int i = 0; i < 3; ++i) Bu[i] = Bx[i * 2] * delta[0] + Bx[i * 2 + 1] * delta[1];
        // exp_delta.w() = MTK::exp(exp_delta.vec(), Bu, scalar(1 / 2)):  1/2 is integer division -> scale 0 -> identity
        Quat q = so3_exp(Bu, 0.0);
        double R[9], A[9], At[9], T1[9], T2[9];
        qtoR(q, R);
        A_matrix(Bu, A);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
        double nR[9];
        for (int i = 0; i < 9; ++i) nR[i] = -R[i];
        mm3(nR, H, T1);
        mm3(T1, At, T2);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j)
                Mx[i * 2 + j] = T2[i * 3 + 0] * Bx[0 * 2 + j] + T2[i * 3 + 1] * Bx[1 * 2 + j] + T2[i * 3 + 2] * Bx[2 * 2 + j];
    }
}

static void state_boxplus(State& s, const double* d) {  // build_manifold.hpp MTK_BOXPLUS over entries
    for (int i = 0; i < 3; ++i) s.pos[i] += d[i];
    s.rot = qmul(s.rot, so3_exp(d + 3, 0.5));
    s.offR = qmul(s.offR, so3_exp(d + 6, 0.5));
    for (int i = 0; i < 3; ++i) s.offT[i] += d[9 + i];
    for (int i = 0; i < 3; ++i) s.vel[i] += d[12 + i];
    for (int i = 0; i < 3; ++i) s.bg[i] += d[15 + i];
    for (int i = 0; i < 3; ++i) s.ba[i] += d[18 + i];
    S2_boxplus(s.grav, d + 21);
}
static void state_boxminus(const State& a, const State& b, double* r) {  // a [-] b
    for (int i = 0; i < 3; ++i) r[i] = a.pos[i] - b.pos[i];
    so3_log(qmul(qconj(b.rot), a.rot), r + 3);
    so3_log(qmul(qconj(b.offR), a.offR), r + 6);
    for (int i = 0; i < 3; ++i) r[9 + i] = a.offT[i] - b.offT[i];
    for (int i = 0; i < 3; ++i) r[12 + i] = a.vel[i] - b.vel[i];
    for (int i = 0; i < 3; ++i) r[15 + i] = a.bg[i] - b.bg[i];
    for (int i = 0; i < 3; ++i) r[18 + i] = a.ba[i] - b.ba[i];
    S2_boxminus(a.grav, b.grav, r + 21);
}

// ------------------------------------------------------------------ LIO front end
struct NearPt { float x, y, z; int32_t ord; };

struct Lio {
    orc_lio_params prm;
    IVoxOracle ivox;
    int nthreads = 1;
    // per-point persistent arrays (laser_mapping.cc:335-339: resized, never cleared)
    std::vector<std::vector<NearPt>> nearest;
    std::vector<float> residuals;
    std::vector<uint8_t> selected;
    std::vector<float> plane;  // 4 per point
    std::vector<float> world;  // scan_down_world_, 3 per point
    // last pass rows
    std::vector<double> h_x, hvec;
    int n_eff = 0;
    double ms_match = 0, ms_jac = 0;

    void resize_point_state(size_t n) {
        nearest.resize(n);
        residuals.resize(n, 0.0f);
        selected.resize(n, 1);
        plane.resize(n * 4, 0.0f);
        world.resize(n * 3);
    }

    // laser_mapping.cc:592-701.  Returns false when ekfom_data.valid=false.
    bool obs_model(const State& s, bool converge, const float* scan, int64_t n, int64_t stride) {
        auto t0 = std::chrono::steady_clock::now();
        // R_wl = (rot * offR).cast<float>() : a float quaternion ; t_wl = (rot*offT + pos).cast<float>()
        Quat qd = qmul(s.rot, s.offR);
        const float qx = (float)qd.x, qy = (float)qd.y, qz = (float)qd.z, qw = (float)qd.w;
        double td[3];
        qrot(s.rot, s.offT, td);
        const float tx = (float)(td[0] + s.pos[0]), ty = (float)(td[1] + s.pos[1]), tz = (float)(td[2] + s.pos[2]);
        const float thr = prm.plane_thr;

#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            const float* pb = (const float*)((const char*)scan + i * stride);
            const float bx = pb[0], by = pb[1], bz = pb[2];
            // Eigen quaternion * vector in float, then + t_wl  (:611-612)
            float uvx = qy * bz - qz * by, uvy = qz * bx - qx * bz, uvz = qx * by - qy * bx;
            uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
            float cx = qy * uvz - qz * uvy, cy = qz * uvx - qx * uvz, cz = qx * uvy - qy * uvx;
            float wx = ((bx + qw * uvx) + cx) + tx;
            float wy = ((by + qw * uvy) + cy) + ty;
            float wz = ((bz + qw * uvz) + cz) + tz;
            world[i * 3 + 0] = wx; world[i * 3 + 1] = wy; world[i * 3 + 2] = wz;

            std::vector<NearPt>& near = nearest[i];
            if (converge) {  // :616-624
                Cand c[5];
                int m = ivox.knn(wx, wy, wz, 5, 5.0, c);
                // Quirk Q4: GetClosestPoint returns before closest_pt.clear() when there are no
                // candidates (ivox3d.h:151-153,199), so a query with an empty stencil keeps the
                // neighbour list its slot held before (previous pass or previous scan).
                if (m > 0) {
                    near.resize(m);
                    for (int k = 0; k < m; ++k) near[k] = NearPt{c[k].p->x, c[k].p->y, c[k].p->z, c[k].p->ord};
                }
                m = (int)near.size();
                bool sel = m >= 3;
                if (sel) {
                    float pts[5][3];
                    for (int k = 0; k < m; ++k) { pts[k][0] = near[k].x; pts[k][1] = near[k].y; pts[k][2] = near[k].z; }
                    sel = esti_plane(&plane[i * 4], pts, m, thr);
                }
                selected[i] = sel ? 1 : 0;
            }
            if (selected[i]) {  // :626-636
                float temp[4] = {wx, wy, wz, 1.0f};
                float pd2 = dot4_sse(&plane[i * 4], temp);
                float bn = std::sqrt((bx * bx + by * by) + bz * bz);
                bool valid_corr = bn > 81 * pd2 * pd2;
                if (valid_corr) {
                    selected[i] = 1;
                    residuals[i] = pd2;
                }
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        ms_match += std::chrono::duration<double, std::milli>(t1 - t0).count();

        // serial stable compaction (:641-655)
        std::vector<int32_t> eff;
        eff.reserve(n);
        for (int64_t i = 0; i < n; ++i)
            if (selected[i]) eff.push_back((int32_t)i);
        n_eff = (int)eff.size();
        if (n_eff < 1) return false;

        h_x.assign((size_t)n_eff * 12, 0.0);
        hvec.assign(n_eff, 0.0);
        double offRd[9], Rd[9];
        qtoR(s.offR, offRd);
        qtoR(s.rot, Rd);
        float off_R[9], Rt[9], off_t[3];
        for (int i = 0; i < 9; ++i) off_R[i] = (float)offRd[i];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = (float)Rd[j * 3 + i];
        for (int i = 0; i < 3; ++i) off_t[i] = (float)s.offT[i];
        const bool ext = prm.extrinsic_est_en != 0;

#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int e = 0; e < n_eff; ++e) {  // :674-698
            int i = eff[e];
            const float* pb = (const float*)((const char*)scan + (int64_t)i * stride);
            const float be[3] = {pb[0], pb[1], pb[2]};
            float pt[3];
            for (int r = 0; r < 3; ++r) pt[r] = ((off_R[r * 3] * be[0] + off_R[r * 3 + 1] * be[1]) + off_R[r * 3 + 2] * be[2]) + off_t[r];
            const float* nv = &plane[i * 4];
            float C[3];
            for (int r = 0; r < 3; ++r) C[r] = (Rt[r * 3] * nv[0] + Rt[r * 3 + 1] * nv[1]) + Rt[r * 3 + 2] * nv[2];
            // A = [pt]x * C  (dense 3x3 * vec with the explicit zeros of SKEW_SYM_MATRIX)
            float A[3];
            A[0] = (0.0f * C[0] + (-pt[2]) * C[1]) + pt[1] * C[2];
            A[1] = (pt[2] * C[0] + 0.0f * C[1]) + (-pt[0]) * C[2];
            A[2] = ((-pt[1]) * C[0] + pt[0] * C[1]) + 0.0f * C[2];
            double* row = &h_x[(size_t)e * 12];
            row[0] = nv[0]; row[1] = nv[1]; row[2] = nv[2];
            row[3] = A[0]; row[4] = A[1]; row[5] = A[2];
            if (ext) {
                // B = ([be]x * off_R^T) * C
                float S[9] = {0.0f, -be[2], be[1], be[2], 0.0f, -be[0], -be[1], be[0], 0.0f};
                float M[9];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        M[r * 3 + c] = (S[r * 3] * off_R[c * 3] + S[r * 3 + 1] * off_R[c * 3 + 1]) + S[r * 3 + 2] * off_R[c * 3 + 2];
                for (int r = 0; r < 3; ++r) row[6 + r] = (M[r * 3] * C[0] + M[r * 3 + 1] * C[1]) + M[r * 3 + 2] * C[2];
                row[9] = C[0]; row[10] = C[1]; row[11] = C[2];
            }
            hvec[e] = -residuals[i];
        }
        auto t2 = std::chrono::steady_clock::now();
        ms_jac += std::chrono::duration<double, std::milli>(t2 - t1).count();
        return true;
    }

    void hth(double* HtH, double* Hth) const {
        for (int i = 0; i < 144; ++i) HtH[i] = 0;
        for (int i = 0; i < 12; ++i) Hth[i] = 0;
        for (int e = 0; e < n_eff; ++e) {
            const double* r = &h_x[(size_t)e * 12];
            for (int a = 0; a < 12; ++a) {
                for (int b = 0; b < 12; ++b) HtH[a * 12 + b] += r[a] * r[b];
                Hth[a] += r[a] * hvec[e];
            }
        }
    }
};

// P projection helpers on row-major 23x23
static const int N = 23;
static inline void rows3_left(double* M, int idx, const double R3[9]) {  // M.block<3,1>(idx,i) = R3 * M.block<3,1>(idx,i)
    for (int i = 0; i < N; ++i) {
        double a = M[(idx)*N + i], b = M[(idx + 1) * N + i], c = M[(idx + 2) * N + i];
        for (int r = 0; r < 3; ++r) M[(idx + r) * N + i] = R3[r * 3] * a + R3[r * 3 + 1] * b + R3[r * 3 + 2] * c;
    }
}
static inline void cols3_right(double* M, int idx, const double R3[9]) {  // M.block<1,3>(i,idx) = M.block<1,3>(i,idx) * R3^T
    for (int i = 0; i < N; ++i) {
        double a = M[i * N + idx], b = M[i * N + idx + 1], c = M[i * N + idx + 2];
        for (int r = 0; r < 3; ++r) M[i * N + idx + r] = a * R3[r * 3] + b * R3[r * 3 + 1] + c * R3[r * 3 + 2];
    }
}
static inline void rows2_left(double* M, int idx, const double R2[4]) {
    for (int i = 0; i < N; ++i) {
        double a = M[idx * N + i], b = M[(idx + 1) * N + i];
        M[idx * N + i] = R2[0] * a + R2[1] * b;
        M[(idx + 1) * N + i] = R2[2] * a + R2[3] * b;
    }
}
static inline void cols2_right(double* M, int idx, const double R2[4]) {
    for (int i = 0; i < N; ++i) {
        double a = M[i * N + idx], b = M[i * N + idx + 1];
        M[i * N + idx] = a * R2[0] + b * R2[1];
        M[i * N + idx + 1] = a * R2[2] + b * R2[3];
    }
}

// esekfom.hpp:1526-1834
static int iekf_update(Lio& L, const float* scan, int64_t n, int64_t stride, double* xio, double* Pio, orc_iekf_stats* st) {
    const int max_iter = L.prm.max_iter;
    const double Rcov = L.prm.R;
    State x = load_state(xio);
    const State x_prop = x;
    std::vector<double> P_prop(Pio, Pio + N * N), P(N * N), Lm(N * N);
    bool valid = true, converge = true;
    int t = 0;
    double K_h[N], K_x[N * N];
    double dx_new[N];
    std::memset(dx_new, 0, sizeof dx_new);
    L.ms_match = L.ms_jac = 0;
    auto tstart = std::chrono::steady_clock::now();
    if (st) { std::memset(st, 0, sizeof *st); }
    int any_valid = 0;
    P = P_prop;  // P_ member starts as the propagated covariance

    for (int i = -1; i < max_iter; ++i) {
        valid = true;
        int pass = st ? st->passes : 0;
        if (st && pass < ORC_MAX_PASSES) {
            store_state(x, st->x_in[pass]);
            st->knn[pass] = converge ? 1 : 0;
        }
        valid = L.obs_model(x, converge, scan, n, stride);
        if (st) {
            if (pass < ORC_MAX_PASSES) {
                st->n_eff[pass] = L.n_eff;
                if (valid) L.hth(st->HtH[pass], st->Hth[pass]);
            }
            st->passes++;
            if (converge) st->knn_passes++;
        }
        if (!valid) continue;
        any_valid = 1;
        const int dof = L.n_eff;
        double dx[N];
        state_boxminus(x, x_prop, dx);
        std::memcpy(dx_new, dx, sizeof dx);
        P = P_prop;
        const int so3_idx[2] = {3, 6};
        for (int k = 0; k < 2; ++k) {  // :1561-1577
            int idx = so3_idx[k];
            double A[9], At[9];
            A_matrix(dx + idx, A);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) At[r * 3 + c] = A[c * 3 + r];
            double a = dx_new[idx], b = dx_new[idx + 1], c = dx_new[idx + 2];
            for (int r = 0; r < 3; ++r) dx_new[idx + r] = At[r * 3] * a + At[r * 3 + 1] * b + At[r * 3 + 2] * c;
            rows3_left(P.data(), idx, At);
            cols3_right(P.data(), idx, At);
        }
        {  // S2 block, idx 21 (:1579-1601)
            int idx = 21;
            double Nx[6], Mx[6], R2[4];
            S2_Nx_yy(x.grav, Nx);
            S2_Mx(x_prop.grav, dx + idx, Mx);
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 2; ++c) R2[r * 2 + c] = Nx[r * 3] * Mx[c] + Nx[r * 3 + 1] * Mx[2 + c] + Nx[r * 3 + 2] * Mx[4 + c];
            double a = dx_new[idx], b = dx_new[idx + 1];
            dx_new[idx] = R2[0] * a + R2[1] * b;
            dx_new[idx + 1] = R2[2] * a + R2[3] * b;
            rows2_left(P.data(), idx, R2);
            cols2_right(P.data(), idx, R2);
        }
        std::memset(K_x, 0, sizeof K_x);
        if (N > dof) {  // :1618-1651  K = P H^T (H P H^T / R + I)^-1 / R
            std::vector<double> Hc((size_t)dof * N, 0.0);
            for (int e = 0; e < dof; ++e)
                for (int c = 0; c < 12; ++c) Hc[(size_t)e * N + c] = L.h_x[(size_t)e * 12 + c];
            std::vector<double> PHt((size_t)N * dof, 0.0), S((size_t)dof * dof, 0.0), Sinv((size_t)dof * dof);
            for (int r = 0; r < N; ++r)
                for (int e = 0; e < dof; ++e) {
                    double s = 0;
                    for (int c = 0; c < N; ++c) s += P[r * N + c] * Hc[(size_t)e * N + c];
                    PHt[(size_t)r * dof + e] = s;
                }
            for (int a = 0; a < dof; ++a)
                for (int b = 0; b < dof; ++b) {
                    double s = 0;
                    for (int c = 0; c < N; ++c) s += Hc[(size_t)a * N + c] * PHt[(size_t)c * dof + b];
                    S[(size_t)a * dof + b] = s / Rcov + (a == b ? 1.0 : 0.0);
                }
            lu_inverse(S.data(), dof, Sinv.data());
            std::vector<double> K((size_t)N * dof);
            for (int r = 0; r < N; ++r)
                for (int e = 0; e < dof; ++e) {
                    double s = 0;
                    for (int c = 0; c < dof; ++c) s += PHt[(size_t)r * dof + c] * Sinv[(size_t)c * dof + e];
                    K[(size_t)r * dof + e] = s / Rcov;
                }
            for (int r = 0; r < N; ++r) {
                double s = 0;
                for (int e = 0; e < dof; ++e) s += K[(size_t)r * dof + e] * L.hvec[e];
                K_h[r] = s;
                for (int c = 0; c < N; ++c) {
                    double s2 = 0;
                    for (int e = 0; e < dof; ++e) s2 += K[(size_t)r * dof + e] * Hc[(size_t)e * N + c];
                    K_x[r * N + c] = s2;
                }
            }
        } else {  // :1685-1713
            double Pr[N * N], P_temp[N * N], P_inv[N * N], HTH[144], HTh[12];
            for (int k = 0; k < N * N; ++k) Pr[k] = P[k] / Rcov;
            lu_inverse(Pr, N, P_temp);
            L.hth(HTH, HTh);
            for (int a = 0; a < 12; ++a)
                for (int b = 0; b < 12; ++b) P_temp[a * N + b] += HTH[a * 12 + b];
            lu_inverse(P_temp, N, P_inv);
            for (int r = 0; r < N; ++r) {
                double s = 0;
                for (int c = 0; c < 12; ++c) s += P_inv[r * N + c] * HTh[c];
                K_h[r] = s;
                for (int c = 0; c < 12; ++c) {
                    double s2 = 0;
                    for (int k = 0; k < 12; ++k) s2 += P_inv[r * N + k] * HTH[k * 12 + c];
                    K_x[r * N + c] = s2;
                }
            }
        }
        double dx_[N];
        for (int r = 0; r < N; ++r) {  // :1719
            double s = 0;
            for (int c = 0; c < N; ++c) s += (K_x[r * N + c] - (r == c ? 1.0 : 0.0)) * dx_new[c];
            dx_[r] = K_h[r] + s;
        }
        state_boxplus(x, dx_);
        converge = true;
        for (int k = 0; k < N; ++k)
            if (std::fabs(dx_[k]) > L.prm.limit[k]) { converge = false; break; }
        if (converge) t++;
        if (!t && i == max_iter - 2) converge = true;

        if (t > 1 || i == max_iter - 1) {  // :1735-1831
            Lm = P;
            for (int k = 0; k < 2; ++k) {
                int idx = so3_idx[k];
                double A[9], At[9];
                A_matrix(dx_ + idx, A);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) At[r * 3 + c] = A[c * 3 + r];
                for (int c = 0; c < N; ++c) {  // L.block<3,1>(idx,i) = res * P.block<3,1>(idx,i)
                    double a = P[idx * N + c], b = P[(idx + 1) * N + c], cc = P[(idx + 2) * N + c];
                    for (int r = 0; r < 3; ++r) Lm[(idx + r) * N + c] = At[r * 3] * a + At[r * 3 + 1] * b + At[r * 3 + 2] * cc;
                }
                for (int c = 0; c < 12; ++c) {  // K_x rows
                    double a = K_x[idx * N + c], b = K_x[(idx + 1) * N + c], cc = K_x[(idx + 2) * N + c];
                    for (int r = 0; r < 3; ++r) K_x[(idx + r) * N + c] = At[r * 3] * a + At[r * 3 + 1] * b + At[r * 3 + 2] * cc;
                }
                cols3_right(Lm.data(), idx, At);
                cols3_right(P.data(), idx, At);
            }
            {
                int idx = 21;
                double Nx[6], Mx[6], R2[4];
                S2_Nx_yy(x.grav, Nx);
                S2_Mx(x_prop.grav, dx_ + idx, Mx);
                for (int r = 0; r < 2; ++r)
                    for (int c = 0; c < 2; ++c) R2[r * 2 + c] = Nx[r * 3] * Mx[c] + Nx[r * 3 + 1] * Mx[2 + c] + Nx[r * 3 + 2] * Mx[4 + c];
                for (int c = 0; c < N; ++c) {
                    double a = P[idx * N + c], b = P[(idx + 1) * N + c];
                    Lm[idx * N + c] = R2[0] * a + R2[1] * b;
                    Lm[(idx + 1) * N + c] = R2[2] * a + R2[3] * b;
                }
                for (int c = 0; c < 12; ++c) {
                    double a = K_x[idx * N + c], b = K_x[(idx + 1) * N + c];
                    K_x[idx * N + c] = R2[0] * a + R2[1] * b;
                    K_x[(idx + 1) * N + c] = R2[2] * a + R2[3] * b;
                }
                cols2_right(Lm.data(), idx, R2);
                cols2_right(P.data(), idx, R2);
            }
            std::vector<double> Pn(N * N);
            for (int r = 0; r < N; ++r)
                for (int c = 0; c < N; ++c) {
                    double s = 0;
                    for (int k = 0; k < 12; ++k) s += K_x[r * N + k] * P[k * N + c];
                    Pn[r * N + c] = Lm[r * N + c] - s;
                }
            P = Pn;
            break;
        }
    }
    store_state(x, xio);
    std::memcpy(Pio, P.data(), sizeof(double) * N * N);
    auto tend = std::chrono::steady_clock::now();
    if (st) {
        st->converged = t > 1;
        st->status = any_valid ? 0 : 1;
        st->ms_match = L.ms_match;
        st->ms_jacobian = L.ms_jac;
        st->ms_solve = std::chrono::duration<double, std::milli>(tend - tstart).count() - L.ms_match - L.ms_jac;
    }
    return any_valid ? 0 : 1;
}

// laser_mapping.cc:525-583
static int64_t map_incremental(Lio& L, const float* scan, int64_t n, int64_t stride, const State& s, bool ekf_inited,
                               int32_t* n_add, int32_t* n_nodown) {
    std::vector<float> to_add, no_down;
    to_add.reserve(n * 3);
    no_down.reserve(n * 3);
    const double fs = L.prm.filter_size_map;
    const float fsf = (float)fs;
    for (int64_t i = 0; i < n; ++i) {
        const float* pb = (const float*)((const char*)scan + i * stride);
        // PointBodyToWorld (:855-864): double quaternion arithmetic, narrowed on store
        double pbd[3] = {pb[0], pb[1], pb[2]}, t1[3], t2[3];
        qrot(s.offR, pbd, t1);
        for (int k = 0; k < 3; ++k) t1[k] = t1[k] + s.offT[k];
        qrot(s.rot, t1, t2);
        float w[3];
        for (int k = 0; k < 3; ++k) w[k] = (float)(t2[k] + s.pos[k]);
        L.world[i * 3 + 0] = w[0]; L.world[i * 3 + 1] = w[1]; L.world[i * 3 + 2] = w[2];
        const std::vector<NearPt>& near = L.nearest[i];
        if (!near.empty() && ekf_inited) {
            float center[3];
            for (int k = 0; k < 3; ++k) center[k] = (std::floor(w[k] / fsf) + 0.5f) * fsf;  // :547-548 (float array expr)
            float d2c[3] = {near[0].x - center[0], near[0].y - center[1], near[0].z - center[2]};
            if (std::fabs((double)d2c[0]) > 0.5 * fs && std::fabs((double)d2c[1]) > 0.5 * fs && std::fabs((double)d2c[2]) > 0.5 * fs) {
                no_down.insert(no_down.end(), w, w + 3);
                continue;
            }
            bool need_add = true;
            float ddx = w[0] - center[0], ddy = w[1] - center[1], ddz = w[2] - center[2];
            float dist = (ddx * ddx + ddy * ddy) + ddz * ddz;
            if (near.size() >= 5) {
                for (int k = 0; k < 5; ++k) {
                    float ex = near[k].x - center[0], ey = near[k].y - center[1], ez = near[k].z - center[2];
                    float dk = (ex * ex + ey * ey) + ez * ez;
                    if ((double)dk < (double)dist + 1e-6) { need_add = false; break; }
                }
            }
            if (need_add) to_add.insert(to_add.end(), w, w + 3);
        } else {
            to_add.insert(to_add.end(), w, w + 3);
        }
    }
    if (n_add) *n_add = (int32_t)(to_add.size() / 3);
    if (n_nodown) *n_nodown = (int32_t)(no_down.size() / 3);
    L.ivox.add_points(to_add.data(), (int64_t)to_add.size() / 3, 12);   // :579
    L.ivox.add_points(no_down.data(), (int64_t)no_down.size() / 3, 12); // :580
    return (int64_t)(to_add.size() + no_down.size()) / 3;
}

}  // namespace orc

using namespace orc;
struct orc_lio { Lio L; };

extern "C" {

orc_lio* orc_lio_create(const orc_lio_params* p) {
    orc_lio* h = new orc_lio();
    h->L.prm = *p;
    h->L.ivox.configure(p->resolution, p->nearby, (size_t)p->capacity_voxels);
    h->L.nthreads = p->num_threads > 0 ? p->num_threads : omp_get_max_threads();
    return h;
}
void orc_lio_destroy(orc_lio* h) { delete h; }
int64_t orc_map_insert(orc_lio* h, const float* xyz, int64_t n, int64_t stride) {
    h->L.ivox.add_points(xyz, n, stride);
    return h->L.ivox.next_ord;
}
int64_t orc_map_num_voxels(orc_lio* h) { return (int64_t)h->L.ivox.grid.size(); }
int64_t orc_map_num_points(orc_lio* h) {
    int64_t s = 0;
    for (auto& nd : h->L.ivox.cache) s += (int64_t)nd.second.size();
    return s;
}
void orc_map_knn5(orc_lio* h, const float* xyz, int64_t n, int64_t stride, int32_t* idx, float* sqdist, int32_t* count) {
#pragma omp parallel for num_threads(h->L.nthreads) schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const float* q = (const float*)((const char*)xyz + i * stride);
        Cand c[5];
        int m = h->L.ivox.knn(q[0], q[1], q[2], 5, 5.0, c);
        for (int k = 0; k < 5; ++k) {
            idx[i * 5 + k] = k < m ? c[k].p->ord : -1;
            sqdist[i * 5 + k] = k < m ? (float)c[k].dist : 0.0f;
        }
        count[i] = m;
    }
}
int64_t orc_map_knn_candidates(orc_lio* h, const float* xyz, int64_t n, int64_t stride) {
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* q = (const float*)((const char*)xyz + i * stride);
        Cand c[5];
        h->L.ivox.knn(q[0], q[1], q[2], 5, 5.0, c, &total);
    }
    return total;
}
int32_t orc_iekf_update(orc_lio* h, const float* scan, int64_t n, int64_t stride, double* x, double* P, orc_iekf_stats* st) {
    h->L.resize_point_state((size_t)n);
    return iekf_update(h->L, scan, n, stride, x, P, st);
}
int32_t orc_obs_model(orc_lio* h, const float* scan, int64_t n, int64_t stride, const double* x, int32_t converge,
                      double* HtH, double* Hth, int32_t* n_eff) {
    h->L.resize_point_state((size_t)n);
    State s = load_state(x);
    bool valid = h->L.obs_model(s, converge != 0, scan, n, stride);
    if (n_eff) *n_eff = h->L.n_eff;
    if (valid && HtH && Hth) h->L.hth(HtH, Hth);
    return valid ? 0 : 1;
}
void orc_point_state(orc_lio* h, int64_t n, float* plane4, float* residual, uint8_t* selected, int32_t* nn_idx5, int32_t* nn_count) {
    Lio& L = h->L;
    for (int64_t i = 0; i < n; ++i) {
        if (plane4) std::memcpy(plane4 + i * 4, &L.plane[i * 4], 16);
        if (residual) residual[i] = L.residuals[i];
        if (selected) selected[i] = L.selected[i];
        if (nn_count) nn_count[i] = (int32_t)L.nearest[i].size();
        if (nn_idx5)
            for (int k = 0; k < 5; ++k) nn_idx5[i * 5 + k] = k < (int)L.nearest[i].size() ? L.nearest[i][k].ord : -1;
    }
}
int32_t orc_last_rows(orc_lio* h, double* h_x, double* hvec, int32_t max_rows) {
    int m = std::min(max_rows, h->L.n_eff);
    std::memcpy(h_x, h->L.h_x.data(), sizeof(double) * 12 * m);
    std::memcpy(hvec, h->L.hvec.data(), sizeof(double) * m);
    return h->L.n_eff;
}
int64_t orc_map_incremental(orc_lio* h, const float* scan, int64_t n, int64_t stride, const double* x, int32_t ekf_inited,
                            int32_t* n_add, int32_t* n_nodown) {
    h->L.resize_point_state((size_t)n);
    State s = load_state(x);
    return map_incremental(h->L, scan, n, stride, s, ekf_inited != 0, n_add, n_nodown);
}
int32_t orc_esti_plane(const float* pts, int32_t n, float thr, float* plane4) {
    float p[5][3];
    for (int i = 0; i < n && i < 5; ++i) { p[i][0] = pts[i * 3]; p[i][1] = pts[i * 3 + 1]; p[i][2] = pts[i * 3 + 2]; }
    return esti_plane(plane4, p, n, thr) ? 1 : 0;
}
void orc_state_boxplus(double* x26, const double* dx23) {
    State s = load_state(x26);
    state_boxplus(s, dx23);
    store_state(s, x26);
}
void orc_state_boxminus(const double* x26, const double* y26, double* dx23) {
    State a = load_state(x26), b = load_state(y26);
    state_boxminus(a, b, dx23);
}
void orc_inverse(const double* A, int32_t n, double* out) { lu_inverse(A, n, out); }

/* esekf::predict (esekfom.hpp:269-374) with the process model of use-ikfom.hpp:36-77 (get_f, df_dx, df_dw), K steps.
 * steps: K x 8 doubles {dt, offs_t, acc_avr[3], angvel_avr[3]} - what ImuProcess::UndistortPcl hands to kf_state.predict for
 * every IMU interval (imu_processing.hpp:190-241); Q12 = diagonal of Q_ (cov_gyr, cov_acc, cov_bias_gyr, cov_bias_acc).
 * poses22 (K x 22, optional) = the IMUpose_ entries pushed after each step: {offs_t, acc_s_last, angvel_last, vel, pos, R}
 * (imu_processing.hpp:225-236).  Quirks kept: the SO3 / S2 blocks of F_x1 are built from MTK::exp(.., scalar(1 / 2)) whose
 * scale is the INTEGER quotient 0, i.e. the identity rotation (esekfom.hpp:307,331). */
void orc_predict(const double* steps, int32_t K, const double* Q12, double* x26, double* P, double* poses22) {
    State s = load_state(x26);
    const int n = 23;
    for (int k = 0; k < K; ++k) {
        const double dt = steps[k * 8], offs_t = steps[k * 8 + 1];
        const double* acc = steps + k * 8 + 2;
        const double* gyr = steps + k * 8 + 5;
        // get_f
        double omega[3], am[3], a_in[3], R[9];
        for (int i = 0; i < 3; ++i) { omega[i] = gyr[i] - s.bg[i]; am[i] = acc[i] - s.ba[i]; }
        qrot(s.rot, am, a_in);
        qtoR(s.rot, R);
        // df_dx (24 x 23, rows by flattened dim) and df_dw (24 x 12): only the non-zero blocks
        double Mx0[6], zero2[2] = {0, 0};
        S2_Mx(s.grav, zero2, Mx0);                 // cov.block<3,2>(12, 21)
        double Hacc[9], RH[9];
        hat3(am, Hacc);
        mm3(R, Hacc, RH);                          // cov.block<3,3>(12, 3) = -R * hat(acc - ba)
        const State before = s;
        // x_.oplus(f_, dt): vect += scale * vec, SO3 *= exp(vec, scale), S2 rotated by exp(0) (build_manifold.hpp MTK_OPLUS)
        for (int i = 0; i < 3; ++i) s.pos[i] += dt * before.vel[i];
        s.rot = qmul(s.rot, so3_exp(omega, dt / 2));
        { const double z3[3] = {0, 0, 0}; s.offR = qmul(s.offR, so3_exp(z3, dt / 2)); }
        for (int i = 0; i < 3; ++i) s.vel[i] += dt * (a_in[i] + before.grav[i]);
        // F_x1 and f_x_final / f_w_final
        double F[23 * 23] = {0}, G[23 * 12] = {0};
        for (int i = 0; i < n; ++i) F[i * n + i] = 1.0;
        double seg[3], A[9];
        for (int i = 0; i < 3; ++i) seg[i] = -1 * omega[i] * dt;
        A_matrix(seg, A);
        // rot rows: f_x_final.block<3,1>(3, i) = A * f_x_.block<3,1>(3, i), f_x_(3..5, 15..17) = -I ; f_w_(3..5, 0..2) = -I
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) { F[(3 + r) * n + 15 + c] += (-A[r * 3 + c]) * dt; G[(3 + r) * 12 + c] = -A[r * 3 + c]; }
        // pos rows: d pos / d vel = I
        for (int r = 0; r < 3; ++r) F[r * n + 12 + r] += 1.0 * dt;
        // vel rows
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) {
                F[(12 + r) * n + 3 + c] += (-RH[r * 3 + c]) * dt;
                F[(12 + r) * n + 18 + c] += (-R[r * 3 + c]) * dt;
                G[(12 + r) * 12 + 3 + c] = -R[r * 3 + c];
            }
            for (int c = 0; c < 2; ++c) F[(12 + r) * n + 21 + c] += Mx0[r * 2 + c] * dt;
        }
        for (int r = 0; r < 3; ++r) { G[(15 + r) * 12 + 6 + r] = 1.0; G[(18 + r) * 12 + 9 + r] = 1.0; }
        // S2 block of F_x1: Nx(x_) * I * Mx(x_before, 0); its f_x rows are zero (df_dx has no grav rows)
        double Nx[6];
        S2_Nx_yy(s.grav, Nx);
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 2; ++c)
                F[(21 + r) * n + 21 + c] = Nx[r * 3] * Mx0[c] + Nx[r * 3 + 1] * Mx0[2 + c] + Nx[r * 3 + 2] * Mx0[4 + c];
        // P_ = F P F^T + (dt G) Q (dt G)^T
        double T[23 * 23], Pn[23 * 23];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double v = 0;
                for (int c = 0; c < n; ++c) v += F[i * n + c] * P[c * n + j];
                T[i * n + j] = v;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double v = 0;
                for (int c = 0; c < n; ++c) v += T[i * n + c] * F[j * n + c];
                double w = 0;
                for (int c = 0; c < 12; ++c) w += ((dt * G[i * 12 + c]) * Q12[c]) * (dt * G[j * 12 + c]);
                Pn[i * n + j] = v + w;
            }
        std::memcpy(P, Pn, sizeof Pn);
        if (poses22) {  // imu_processing.hpp:225-236
            double* o = poses22 + (size_t)k * 22;
            double am2[3], as[3], R2[9];
            for (int i = 0; i < 3; ++i) am2[i] = acc[i] - s.ba[i];
            qrot(s.rot, am2, as);
            qtoR(s.rot, R2);
            o[0] = offs_t;
            for (int i = 0; i < 3; ++i) { o[1 + i] = as[i] + s.grav[i]; o[4 + i] = gyr[i] - s.bg[i]; o[7 + i] = s.vel[i]; o[10 + i] = s.pos[i]; }
            for (int i = 0; i < 9; ++i) o[13 + i] = R2[i];
        }
    }
    store_state(s, x26);
}

}  // extern "C"
