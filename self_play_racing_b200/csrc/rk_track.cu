// Track pool construction on the device.
//
// Replaces Track.__init__ and friends (reference environment/track.py:61-148)
// and gen_tracks/gen_random_track (track.py:4-56).  One CTA builds one track:
// thread 0 solves the two periodic cubic splines (the condensed-tridiagonal
// algorithm of scipy.interpolate.CubicSpline(bc_type='periodic'),
// scipy/interpolate/_cubic.py, which track.py:106-107 calls), then the CTA
// evaluates N waypoints, normals, boundaries, the segment table and the
// bounding-circle chunks used by the culled query path.
//
// All float64 arithmetic that decides discrete events downstream uses the
// round-to-nearest intrinsics (no FMA contraction) in numpy's operation order,
// so tables built from the reference's own waypoints are bit-identical to the
// reference's tables.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "rk_types.cuh"

namespace rk {

namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct BuildArgs {
    TrackMeta* meta;
    const double* ctrl_xy;  // concatenated (x, y) control points, or nullptr
    int from_waypoints;     // waypoints already stored in wx/wy
    double *wx, *wy, *nx, *ny, *sx, *sy, *v2x, *v2y;
    float2 *wpt, *bpt;
    float4 *wchunk, *bchunk;
};

// Periodic cubic spline through (t[i], y[i]), i = 0..n-1, y[n-1] == y[0].
// Writes PPoly coefficients c[k][i] (k = 0 highest power) for the n-1 intervals.
// Follows scipy _cubic.py (CubicSpline.__init__ periodic branch and
// CubicHermiteSpline.__init__); the (n-2)x(n-2) systems are solved the way
// LAPACK dgtsv does for a diagonally dominant matrix (no row interchange).
__device__ void periodic_spline(int n, const double* t, const double* y, double* c0, double* c1, double* c2,
                                double* c3, double* dx, double* slope, double* s, double* w0, double* w1,
                                double* w2) {
    const int m = n - 1;  // intervals; unknowns s[0..m-1]
    for (int i = 0; i < m; ++i) {
        dx[i] = dsub(t[i + 1], t[i]);
        slope[i] = ddiv(dsub(y[i + 1], y[i]), dx[i]);
    }
    // rhs b[0..m-1]
    double* b = w0;
    for (int i = 1; i < m; ++i)
        b[i] = dmul(3.0, dadd(dmul(dx[i], slope[i - 1]), dmul(dx[i - 1], slope[i])));
    b[0] = dmul(3.0, dadd(dmul(dx[0], slope[m - 1]), dmul(dx[m - 1], slope[0])));
    // condensed (m-1)x(m-1) tridiagonal: diag d, upper du, lower dl
    const int q = m - 1;
    double* d1 = w1;   // working diagonal for rhs b1
    double* s2 = w2;   // second rhs / solution
    // matrix rows: row 0: diag 2(dx[m-1]+dx[0]), upper dx[m-1]
    //              row i: lower dx[i], diag 2(dx[i-1]+dx[i]), upper dx[i-1]
    for (int i = 0; i < q; ++i) {
        d1[i] = (i == 0) ? dmul(2.0, dadd(dx[m - 1], dx[0])) : dmul(2.0, dadd(dx[i - 1], dx[i]));
        s[i] = b[i];
        s2[i] = 0.0;
    }
    s2[0] = -dx[0];
    s2[q - 1] = -dx[m - 3 >= 0 ? m - 3 : 0];  // a_m2_m1 = dx[-3]
    // forward elimination (shared factorisation for both right-hand sides)
    for (int i = 0; i < q - 1; ++i) {
        const double dl = dx[i + 1];                        // matrix[i+1][i]
        const double du = (i == 0) ? dx[m - 1] : dx[i - 1];  // matrix[i][i+1]
        const double fact = ddiv(dl, d1[i]);
        d1[i + 1] = dsub(d1[i + 1], dmul(fact, du));
        s[i + 1] = dsub(s[i + 1], dmul(fact, s[i]));
        s2[i + 1] = dsub(s2[i + 1], dmul(fact, s2[i]));
    }
    // back substitution
    s[q - 1] = ddiv(s[q - 1], d1[q - 1]);
    s2[q - 1] = ddiv(s2[q - 1], d1[q - 1]);
    for (int i = q - 2; i >= 0; --i) {
        const double du = (i == 0) ? dx[m - 1] : dx[i - 1];
        s[i] = ddiv(dsub(s[i], dmul(du, s[i + 1])), d1[i]);
        s2[i] = ddiv(dsub(s2[i], dmul(du, s2[i + 1])), d1[i]);
    }
    // last unknown, then the full solution (scipy: s_m1, s[:-2] = s1 + s_m1*s2)
    const double a_m1_0 = dx[m - 2], a_m1_m2 = dx[m - 1];
    const double a_m1_m1 = dmul(2.0, dadd(dx[m - 1], dx[m - 2]));
    const double bl = dmul(3.0, dadd(dmul(dx[m - 1], slope[m - 2]), dmul(dx[m - 2], slope[m - 1])));
    const double num = dsub(dsub(bl, dmul(a_m1_0, s[0])), dmul(a_m1_m2, s[q - 1]));
    const double den = dadd(dadd(a_m1_m1, dmul(a_m1_0, s2[0])), dmul(a_m1_m2, s2[q - 1]));
    const double s_m1 = ddiv(num, den);
    for (int i = 0; i < q; ++i) s[i] = dadd(s[i], dmul(s_m1, s2[i]));
    s[m - 1] = s_m1;
    s[m] = s[0];
    for (int i = 0; i < m; ++i) {
        const double tt = ddiv(dsub(dadd(s[i], s[i + 1]), dmul(2.0, slope[i])), dx[i]);
        c0[i] = ddiv(tt, dx[i]);
        c1[i] = dsub(ddiv(dsub(slope[i], s[i]), dx[i]), tt);
        c2[i] = s[i];
        c3[i] = y[i];
    }
}

__global__ void __launch_bounds__(256) build_track_kernel(BuildArgs a) {
    __shared__ double knots[kMaxKnots], px[kMaxKnots], py[kMaxKnots];
    __shared__ double cx[4][kMaxKnots], cy[4][kMaxKnots];
    __shared__ double wk[6][kMaxKnots];
    __shared__ double red[4][256];
    __shared__ double s_org[2];

    TrackMeta& tm = a.meta[blockIdx.x];
    const int N = tm.n_wp, off = tm.wp_off, tid = threadIdx.x, nt = blockDim.x;
    double* wx = a.wx + off;
    double* wy = a.wy + off;
    double* nx = a.nx + off;
    double* ny = a.ny + off;

    if (!a.from_waypoints) {
        const int nc = tm.n_ctrl, n = nc + 1;
        if (tid == 0) {
            const double* cp = a.ctrl_xy + 2 * (size_t)tm.ctrl_off;
            for (int i = 0; i < n; ++i) {
                px[i] = cp[2 * (i % nc)];
                py[i] = cp[2 * (i % nc) + 1];
            }
            knots[0] = 0.0;  // track.py:105 chord-length parameter
            for (int i = 1; i < n; ++i) {
                const double ddx = dsub(px[i], px[i - 1]), ddy = dsub(py[i], py[i - 1]);
                knots[i] = dadd(knots[i - 1], sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy))));
            }
            periodic_spline(n, knots, px, cx[0], cx[1], cx[2], cx[3], wk[0], wk[1], wk[2], wk[3], wk[4], wk[5]);
            periodic_spline(n, knots, py, cy[0], cy[1], cy[2], cy[3], wk[0], wk[1], wk[2], wk[3], wk[4], wk[5]);
        }
        __syncthreads();
        // track.py:110-114: N samples of linspace(0, t_end, N, endpoint=False)
        const double step = ddiv(knots[n - 1], (double)N);
        for (int k = tid; k < N; k += nt) {
            const double tw = dmul((double)k, step);
            int i = 0;
            while (i < n - 2 && tw >= knots[i + 1]) ++i;
            const double s = dsub(tw, knots[i]);
            // scipy _ppoly evaluate_poly1: res = sum c[k] * s^p accumulated from p = 0
            const double s2 = dmul(s, s), s3 = dmul(s2, s);
            wx[k] = dadd(dadd(dadd(cx[3][i], dmul(cx[2][i], s)), dmul(cx[1][i], s2)), dmul(cx[0][i], s3));
            wy[k] = dadd(dadd(dadd(cy[3][i], dmul(cy[2][i], s)), dmul(cy[1][i], s2)), dmul(cy[0][i], s3));
        }
    }
    __syncthreads();

    // bounding box (track.py:82-91) and fp32 origin
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int k = tid; k < N; k += nt) {
        mnx = fmin(mnx, wx[k]); mxx = fmax(mxx, wx[k]);
        mny = fmin(mny, wy[k]); mxy = fmax(mxy, wy[k]);
    }
    red[0][tid] = mnx; red[1][tid] = mxx; red[2][tid] = mny; red[3][tid] = mxy;
    __syncthreads();
    for (int s = nt / 2; s > 0; s >>= 1) {
        if (tid < s) {
            red[0][tid] = fmin(red[0][tid], red[0][tid + s]);
            red[1][tid] = fmax(red[1][tid], red[1][tid + s]);
            red[2][tid] = fmin(red[2][tid], red[2][tid + s]);
            red[3][tid] = fmax(red[3][tid], red[3][tid + s]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double ex = dsub(red[1][0], red[0][0]), ey = dsub(red[3][0], red[2][0]);
        tm.max_track_distance = sqrt(dadd(dmul(ex, ex), dmul(ey, ey)));
        s_org[0] = 0.5 * (red[0][0] + red[1][0]);
        s_org[1] = 0.5 * (red[2][0] + red[3][0]);
        tm.org_x = s_org[0];
        tm.org_y = s_org[1];
        tm.start_x = wx[0];
        tm.start_y = wy[0];
        tm.start_angle = atan2(dsub(wy[1], wy[0]), dsub(wx[1], wx[0]));  // track.py:154-157
    }
    // normals (track.py:117-124)
    for (int k = tid; k < N; k += nt) {
        const int k1 = (k + 1 == N) ? 0 : k + 1;
        double tx = dsub(wx[k1], wx[k]), ty = dsub(wy[k1], wy[k]);
        double len = sqrt(dadd(dmul(tx, tx), dmul(ty, ty)));
        if (len == 0.0) len = 1.0;
        tx = ddiv(tx, len);
        ty = ddiv(ty, len);
        nx[k] = -ty;
        ny[k] = tx;
    }
    __syncthreads();
    if (tid == 0) {
        tm.start_nx = nx[0];
        tm.start_ny = ny[0];
    }
    // boundaries and the segment table (track.py:93-94,126-148): left side first
    const double w = tm.width, ox = s_org[0], oy = s_org[1];
    double* sx = a.sx + 2 * (size_t)off;
    double* sy = a.sy + 2 * (size_t)off;
    double* v2x = a.v2x + 2 * (size_t)off;
    double* v2y = a.v2y + 2 * (size_t)off;
    float2* bpt = a.bpt + tm.bpt_off;
    float2* wpt = a.wpt + off;
    for (int k = tid; k < N; k += nt) {
        const int k1 = (k + 1 == N) ? 0 : k + 1;
        const double lx = dadd(wx[k], dmul(nx[k], w)), ly = dadd(wy[k], dmul(ny[k], w));
        const double lx1 = dadd(wx[k1], dmul(nx[k1], w)), ly1 = dadd(wy[k1], dmul(ny[k1], w));
        const double rx = dsub(wx[k], dmul(nx[k], w)), ry = dsub(wy[k], dmul(ny[k], w));
        const double rx1 = dsub(wx[k1], dmul(nx[k1], w)), ry1 = dsub(wy[k1], dmul(ny[k1], w));
        sx[k] = lx; sy[k] = ly; v2x[k] = dsub(lx1, lx); v2y[k] = dsub(ly1, ly);
        sx[N + k] = rx; sy[N + k] = ry; v2x[N + k] = dsub(rx1, rx); v2y[N + k] = dsub(ry1, ry);
        bpt[k] = make_float2((float)(lx - ox), (float)(ly - oy));
        bpt[N + 1 + k] = make_float2((float)(rx - ox), (float)(ry - oy));
        if (k == 0) {  // each row is closed: point N repeats point 0
            bpt[N] = bpt[0];
            bpt[2 * N + 1] = bpt[N + 1];
        }
        wpt[k] = make_float2((float)(wx[k] - ox), (float)(wy[k] - oy));
    }
    __syncthreads();
    // bounding circles.  Waypoint chunk c covers waypoints [c*kChunk, (c+1)*kChunk);
    // boundary chunk c of a side covers segments [c*kRaySegs, (c+1)*kRaySegs),
    // i.e. points c*kRaySegs .. end inclusive (point N = point 0).
    const int nwc = (N + kChunk - 1) / kChunk, nrc = (N + kRaySegs - 1) / kRaySegs;
    for (int c = tid; c < nwc + 2 * nrc; c += nt) {
        const bool is_wp = c < nwc;
        const int kind = is_wp ? 2 : (c - nwc) / nrc;          // 0 left, 1 right, 2 waypoints
        const int cc = is_wp ? c : (c - nwc) % nrc;
        const int span = is_wp ? kChunk : kRaySegs;
        const int k0 = cc * span;
        const int k1 = min(k0 + span, N);
        const int last = is_wp ? k1 - 1 : k1;  // segments need their end point too
        double bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY;
        for (int k = k0; k <= last; ++k) {
            const int kk = (k == N) ? 0 : k;
            const double qx = is_wp ? wx[kk] : sx[kind * N + kk];
            const double qy = is_wp ? wy[kk] : sy[kind * N + kk];
            bx0 = fmin(bx0, qx); bx1 = fmax(bx1, qx);
            by0 = fmin(by0, qy); by1 = fmax(by1, qy);
        }
        const double ccx = 0.5 * (bx0 + bx1), ccy = 0.5 * (by0 + by1);
        double r2 = 0.0;
        for (int k = k0; k <= last; ++k) {
            const int kk = (k == N) ? 0 : k;
            const double qx = is_wp ? wx[kk] : sx[kind * N + kk];
            const double qy = is_wp ? wy[kk] : sy[kind * N + kk];
            r2 = fmax(r2, (qx - ccx) * (qx - ccx) + (qy - ccy) * (qy - ccy));
        }
        // margin covers fp32 rounding of the tables, the centre and the query
        const float4 circ = make_float4((float)(ccx - ox), (float)(ccy - oy), (float)(sqrt(r2) + 2e-3), 0.f);
        if (is_wp)
            a.wchunk[tm.wchunk_off + cc] = circ;
        else
            a.bchunk[tm.bchunk_off + kind * nrc + cc] = circ;
    }
}

// gen_tracks + gen_random_track (track.py:4-56) with Philox draws instead of
// the global MT19937 stream: same distributions, not the same numbers.
__global__ void gen_control_points_kernel(uint64_t seed, int n_tracks, int max_ctrl, double* ctrl_xy, int32_t* n_ctrl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    uint32_t ctr = 0;
    auto draw = [&]() {  // uniform double in (0, 1)
        uint32_t c[4] = {ctr++, (uint32_t)t, 0x7261636bu, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        return u01d(c[0], c[1]);
    };
    const int n = 10 + min(4, (int)(draw() * 5.0));             // randint(10, 15)
    const int base = 50 + min(29, (int)(draw() * 30.0));        // randint(50, 80)
    const int vhi = base / 2 - 10;                              // randint(10, base//2 - 10)
    const int var = 10 + min(vhi - 11, (int)(draw() * (vhi - 10)));
    const double jitter = 0.2 + 0.5 * draw();                   // uniform(0.2, 0.7)
    const double smooth = 0.2 + 0.5 * draw();
    double ang[16], rad[16];
    const double two_pi = 2.0 * M_PI, spacing = two_pi / n;
    for (int i = 0; i < n; ++i) {
        const double off = (2.0 * draw() - 1.0) * (jitter * spacing / 2.0);
        double a = fmod(i * spacing + off, two_pi);
        if (a < 0) a += two_pi;
        ang[i] = a;
    }
    for (int i = 1; i < n; ++i) {  // np.sort
        const double v = ang[i];
        int j = i - 1;
        while (j >= 0 && ang[j] > v) { ang[j + 1] = ang[j]; --j; }
        ang[j + 1] = v;
    }
    for (int i = 0; i < n; ++i) {
        const double raw = base + (2.0 * draw() - 1.0) * var;
        rad[i] = (i == 0) ? raw : (1.0 - smooth) * raw + smooth * rad[i - 1];
    }
    rad[0] = (rad[0] + rad[n - 1]) / 2.0;
    n_ctrl[t] = n;
    double* out = ctrl_xy + 2 * (size_t)t * max_ctrl;
    for (int i = 0; i < n; ++i) {
        out[2 * i] = rad[i] * cos(ang[i]);
        out[2 * i + 1] = rad[i] * sin(ang[i]);
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct PoolBuffers {
    int n_tracks = 0;
    size_t total_wp = 0, total_chunks = 0, total_ctrl = 0;
    TrackMeta* meta = nullptr;
    int32_t* env_to_track = nullptr;
    double *wx = nullptr, *wy = nullptr, *nx = nullptr, *ny = nullptr;
    double *sx = nullptr, *sy = nullptr, *v2x = nullptr, *v2y = nullptr;
    double* ctrl = nullptr;
    float2 *wpt = nullptr, *bpt = nullptr;
    float4 *wchunk = nullptr, *bchunk = nullptr;
    float4* bseg = nullptr;
    uint32_t* gcell = nullptr;
    uint16_t* glist = nullptr;
    std::vector<TrackMeta> host_meta;

    void release() {
        void* ptrs[] = {meta, env_to_track, wx, wy, nx, ny, sx, sy, v2x, v2y, ctrl, wpt, bpt, wchunk, bchunk, bseg, gcell, glist};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        *this = PoolBuffers();
    }
    TrackPool view() const {
        TrackPool v;
        v.meta = meta; v.env_to_track = env_to_track;
        v.wx = wx; v.wy = wy; v.nx = nx; v.ny = ny;
        v.sx = sx; v.sy = sy; v.v2x = v2x; v.v2y = v2y;
        v.wpt = wpt; v.bpt = bpt; v.wchunk = wchunk; v.bchunk = bchunk;
        v.bseg = bseg; v.gcell = gcell; v.glist = glist;
        return v;
    }
};

#define RK_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(err, errlen, "%s failed: %s", #call, cudaGetErrorString(e_));        \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

// Uniform grids over the boundary segments (RK_QUERY_GRID), built on the host from the fp32 boundary rows the
// device has just produced -- a one-off at pool creation.  Cell size RK_B200_GRID_CELL (default 3 track units).
// A segment is listed in every cell that its bounding box, grown by kGridMargin, overlaps: a ray that crosses the
// segment inside a cell (or within fp32 rounding of it) therefore finds it in that cell's list.  The grid covers the
// boundary's bounding box plus 8 units, so every car position that can occur (cars stop at the wall) lies inside; the
// kernel falls back to the exact scan for origins outside.
static int build_grids(PoolBuffers& pb, size_t bpts, char* err, size_t errlen) {
    float cell = 3.0f;
    if (const char* env = getenv("RK_B200_GRID_CELL")) cell = (float)atof(env);
    if (!(cell >= 0.5f && cell <= 64.f)) cell = 3.0f;
    std::vector<float2> pts(bpts);
    RK_CUDA(cudaMemcpy(pts.data(), pb.bpt, bpts * sizeof(float2), cudaMemcpyDeviceToHost));
    std::vector<float4> segs(2 * pb.total_wp);
    std::vector<uint32_t> cells;
    std::vector<uint16_t> list;
    std::vector<uint32_t> count, fill;
    const float pad = 8.0f;
    for (int t = 0; t < pb.n_tracks; ++t) {
        TrackMeta& m = pb.host_meta[t];
        const int N = m.n_wp;
        if (2 * N > 65535) {
            snprintf(err, errlen, "track %d has %d waypoints (segment ids are 16-bit: at most 32767)", t, N);
            return 1;
        }
        const float2* row[2] = {pts.data() + m.bpt_off, pts.data() + m.bpt_off + (N + 1)};
        float x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
        for (int sd = 0; sd < 2; ++sd)
            for (int k = 0; k <= N; ++k) {
                x0 = fminf(x0, row[sd][k].x); x1 = fmaxf(x1, row[sd][k].x);
                y0 = fminf(y0, row[sd][k].y); y1 = fmaxf(y1, row[sd][k].y);
            }
        m.gx0 = x0 - pad; m.gy0 = y0 - pad; m.gcell = cell; m.ginv = 1.0f / cell;
        m.gnx = (int)ceilf((x1 + pad - m.gx0) / cell) + 1;
        m.gny = (int)ceilf((y1 + pad - m.gy0) / cell) + 1;
        m.gcell_off = (int32_t)cells.size();
        m.glist_off = (int32_t)list.size();
        const size_t ncell = (size_t)m.gnx * m.gny;
        if (cells.size() + ncell > 0x7fffffffu) {
            snprintf(err, errlen, "track pool too large for the ray grid (cells)");
            return 1;
        }
        count.assign(ncell, 0);
        auto range = [&](float a, float b, float org, int n, int& lo, int& hi) {
            lo = (int)floorf((fminf(a, b) - kGridMargin - org) * m.ginv);
            hi = (int)floorf((fmaxf(a, b) + kGridMargin - org) * m.ginv);
            lo = lo < 0 ? 0 : lo; hi = hi > n - 1 ? n - 1 : hi;
        };
        for (int pass = 0; pass < 2; ++pass) {
            for (int sg = 0; sg < 2 * N; ++sg) {
                const int sd = sg >= N, k = sg - sd * N;
                const float2 p = row[sd][k], q = row[sd][k + 1];
                if (pass == 0) segs[2 * (size_t)m.wp_off + sg] = make_float4(p.x, p.y, q.x - p.x, q.y - p.y);
                int cx0, cx1, cy0, cy1;
                range(p.x, q.x, m.gx0, m.gnx, cx0, cx1);
                range(p.y, q.y, m.gy0, m.gny, cy0, cy1);
                for (int cy = cy0; cy <= cy1; ++cy)
                    for (int cx = cx0; cx <= cx1; ++cx) {
                        const size_t c = (size_t)cy * m.gnx + cx;
                        if (pass == 0) ++count[c];
                        else list[m.glist_off + fill[c]++] = (uint16_t)sg;
                    }
            }
            if (pass == 0) {
                fill.assign(ncell, 0);
                uint32_t run = 0;
                for (size_t c = 0; c < ncell; ++c) {
                    if (count[c] > 1023 || run >= (1u << 22)) {
                        snprintf(err, errlen, "track %d is too dense for the ray grid (%u segments in one %.1f-unit cell)", t, count[c], cell);
                        return 1;
                    }
                    cells.push_back((run << 10) | count[c]);
                    fill[c] = run;
                    run += count[c];
                }
                list.resize(list.size() + run);
            }
        }
    }
    RK_CUDA(cudaMalloc(&pb.bseg, segs.size() * sizeof(float4)));
    RK_CUDA(cudaMalloc(&pb.gcell, (cells.size() + 1) * sizeof(uint32_t)));
    RK_CUDA(cudaMalloc(&pb.glist, (list.size() + 1) * sizeof(uint16_t)));
    RK_CUDA(cudaMemcpy(pb.bseg, segs.data(), segs.size() * sizeof(float4), cudaMemcpyHostToDevice));
    RK_CUDA(cudaMemcpy(pb.gcell, cells.data(), cells.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RK_CUDA(cudaMemcpy(pb.glist, list.data(), list.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    RK_CUDA(cudaMemcpy(pb.meta, pb.host_meta.data(), pb.n_tracks * sizeof(TrackMeta), cudaMemcpyHostToDevice));
    return 0;
}

// n_wp[t] waypoints per track; exactly one of (dev_ctrl != nullptr) or
// (host_wp != nullptr) supplies the geometry.
int build_pool(PoolBuffers& pb, int n_tracks, const int32_t* n_ctrl, const int32_t* ctrl_off, size_t total_ctrl,
               const double* host_ctrl, const int32_t* n_wp, const double* host_wp, const double* widths,
               const int32_t* host_env_to_track, int E, char* err, size_t errlen) {
    pb.release();
    if (n_tracks <= 0) {
        snprintf(err, errlen, "n_tracks must be positive");
        return 1;
    }
    pb.n_tracks = n_tracks;
    pb.host_meta.assign(n_tracks, TrackMeta());
    size_t wp = 0, ch = 0, bch = 0, bpts = 0;
    for (int t = 0; t < n_tracks; ++t) {
        TrackMeta& m = pb.host_meta[t];
        if (n_wp[t] < 4) {
            snprintf(err, errlen, "track %d has %d waypoints (need >= 4)", t, n_wp[t]);
            return 1;
        }
        if (host_ctrl && (n_ctrl[t] < 3 || n_ctrl[t] + 1 > kMaxKnots)) {
            snprintf(err, errlen, "track %d has %d control points (supported: 3..%d)", t, n_ctrl[t], kMaxKnots - 1);
            return 1;
        }
        const int nch = (n_wp[t] + kChunk - 1) / kChunk;
        const int nrc = (n_wp[t] + kRaySegs - 1) / kRaySegs;
        m.n_wp = n_wp[t];
        m.wp_off = (int32_t)wp;
        m.n_wchunk = nch;
        m.wchunk_off = (int32_t)ch;
        m.n_bchunk = 2 * nrc;
        m.bchunk_off = (int32_t)bch;
        m.bpt_off = (int32_t)bpts;
        m.n_ctrl = n_ctrl ? n_ctrl[t] : 0;
        m.ctrl_off = ctrl_off ? ctrl_off[t] : 0;
        m.width = widths[t];
        wp += n_wp[t];
        ch += nch;
        bch += 2 * nrc;
        bpts += 2 * ((size_t)n_wp[t] + 1);
    }
    pb.total_wp = wp;
    pb.total_chunks = ch;
    pb.total_ctrl = total_ctrl;
    RK_CUDA(cudaMalloc(&pb.meta, n_tracks * sizeof(TrackMeta)));
    RK_CUDA(cudaMalloc(&pb.env_to_track, (size_t)E * sizeof(int32_t)));
    double** d8[] = {&pb.wx, &pb.wy, &pb.nx, &pb.ny};
    for (double** p : d8) RK_CUDA(cudaMalloc(p, wp * sizeof(double)));
    double** d16[] = {&pb.sx, &pb.sy, &pb.v2x, &pb.v2y};
    for (double** p : d16) RK_CUDA(cudaMalloc(p, 2 * wp * sizeof(double)));
    RK_CUDA(cudaMalloc(&pb.wpt, wp * sizeof(float2)));
    RK_CUDA(cudaMalloc(&pb.bpt, bpts * sizeof(float2)));
    RK_CUDA(cudaMalloc(&pb.wchunk, ch * sizeof(float4)));
    RK_CUDA(cudaMalloc(&pb.bchunk, bch * sizeof(float4)));
    RK_CUDA(cudaMemcpy(pb.meta, pb.host_meta.data(), n_tracks * sizeof(TrackMeta), cudaMemcpyHostToDevice));
    std::vector<int32_t> e2t(E);
    for (int e = 0; e < E; ++e) {
        const int t = host_env_to_track ? host_env_to_track[e] : e % n_tracks;
        if (t < 0 || t >= n_tracks) {
            snprintf(err, errlen, "env_to_track[%d] = %d out of range", e, t);
            return 1;
        }
        e2t[e] = t;
    }
    RK_CUDA(cudaMemcpy(pb.env_to_track, e2t.data(), (size_t)E * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (host_ctrl) {
        RK_CUDA(cudaMalloc(&pb.ctrl, 2 * total_ctrl * sizeof(double)));
        RK_CUDA(cudaMemcpy(pb.ctrl, host_ctrl, 2 * total_ctrl * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (host_wp) {  // de-interleave (x, y) pairs
        std::vector<double> hx(wp), hy(wp);
        for (size_t i = 0; i < wp; ++i) {
            hx[i] = host_wp[2 * i];
            hy[i] = host_wp[2 * i + 1];
        }
        RK_CUDA(cudaMemcpy(pb.wx, hx.data(), wp * sizeof(double), cudaMemcpyHostToDevice));
        RK_CUDA(cudaMemcpy(pb.wy, hy.data(), wp * sizeof(double), cudaMemcpyHostToDevice));
    }
    BuildArgs a;
    a.meta = pb.meta;
    a.ctrl_xy = pb.ctrl;
    a.from_waypoints = host_wp ? 1 : 0;
    a.wx = pb.wx; a.wy = pb.wy; a.nx = pb.nx; a.ny = pb.ny;
    a.sx = pb.sx; a.sy = pb.sy; a.v2x = pb.v2x; a.v2y = pb.v2y;
    a.wpt = pb.wpt; a.bpt = pb.bpt; a.wchunk = pb.wchunk; a.bchunk = pb.bchunk;
    build_track_kernel<<<n_tracks, 256>>>(a);
    count_launch();
    RK_CUDA(cudaGetLastError());
    RK_CUDA(cudaDeviceSynchronize());
    RK_CUDA(cudaMemcpy(pb.host_meta.data(), pb.meta, n_tracks * sizeof(TrackMeta), cudaMemcpyDeviceToHost));
    return build_grids(pb, bpts, err, errlen);
}

// Procedural control points on the device, returned to the host so that the
// common build path can size its tables (a one-off at pool creation).
int generate_control_points(uint64_t seed, int n_tracks, std::vector<double>& ctrl, std::vector<int32_t>& n_ctrl,
                            char* err, size_t errlen) {
    const int max_ctrl = 16;
    double* d_ctrl = nullptr;
    int32_t* d_n = nullptr;
    RK_CUDA(cudaMalloc(&d_ctrl, (size_t)n_tracks * max_ctrl * 2 * sizeof(double)));
    RK_CUDA(cudaMalloc(&d_n, (size_t)n_tracks * sizeof(int32_t)));
    gen_control_points_kernel<<<(n_tracks + 127) / 128, 128>>>(seed, n_tracks, max_ctrl, d_ctrl, d_n);
    count_launch();
    RK_CUDA(cudaGetLastError());
    std::vector<double> padded((size_t)n_tracks * max_ctrl * 2);
    n_ctrl.resize(n_tracks);
    RK_CUDA(cudaMemcpy(padded.data(), d_ctrl, padded.size() * sizeof(double), cudaMemcpyDeviceToHost));
    RK_CUDA(cudaMemcpy(n_ctrl.data(), d_n, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyDeviceToHost));
    cudaFree(d_ctrl);
    cudaFree(d_n);
    ctrl.clear();
    for (int t = 0; t < n_tracks; ++t)
        ctrl.insert(ctrl.end(), padded.begin() + (size_t)t * max_ctrl * 2,
                    padded.begin() + (size_t)t * max_ctrl * 2 + 2 * n_ctrl[t]);
    return 0;
}

PoolBuffers* pool_new() { return new PoolBuffers(); }
void pool_delete(PoolBuffers* p) {
    if (p) {
        p->release();
        delete p;
    }
}
TrackPool pool_view(const PoolBuffers* p) { return p->view(); }
int pool_num_tracks(const PoolBuffers* p) { return p ? p->n_tracks : 0; }
const TrackMeta* pool_host_meta(const PoolBuffers* p, int t) { return &p->host_meta[t]; }
const double* pool_ctrl(const PoolBuffers* p) { return p->ctrl; }

}  // namespace rk
