// The fused environment step.
//
// Replaces, in one launch for E environments (reference paths relative to its
// root): Car.update (environment/car.py:45-80), Track.closest_waypoint_idx /
// check_collision / raycast (environment/track.py:150-198),
// MultiTrack.raycast_with_cars (environment/multi_track.py:5-44),
// MultiCar.rectangles_intersect (environment/multi_car.py:16-43),
// RacingEnv.step/_get_obs (environment/racing_env.py:44-166),
// MultiRacingEnv.step/calc_reward/place/_get_obs
// (environment/multi_racing_env.py:48-268) and gymnasium's NEXT_STEP auto-reset
// plus RecordEpisodeStatistics (call sites agent/ppo.py:70,88,114-130).
//
// Thread mapping.  A warp owns EPW = 32/A consecutive environments.  In the
// scalar phases (dynamics, wall test, SAT, reward, termination, placement,
// reset, non-ray observation slots) every lane is ONE CAR (lane = g*A + a), so
// those phases run at full warp width.  For the two geometric queries (waypoint
// argmin, raycast) the warp walks over its environments and all 32 lanes
// cooperate on one query set, exchanging data through shared memory and
// reducing with warp shuffles.  No block-level barrier is used.
//
// Arithmetic contract.  The car state and everything that decides a discrete
// event (waypoint argmin, wall test, SAT, checkpoints, finish, placement) is
// float64 in the reference's operation order with round-to-nearest intrinsics
// (no FMA contraction), so those events are bit-comparable with the reference.
// RK_QUERY_CULLED finds argmin / ray candidates in fp32 over bounding-circle
// chunks and re-evaluates every winner in float64 with the reference's formula.
#include <math.h>
#include <stdio.h>

#include "rk_types.cuh"

namespace rk {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kDt = 0.05;            // car.py:45
constexpr double kMaxSpeed = 30.0;      // car.py:4
constexpr double kTwoPi = 6.283185307179586;
constexpr double kMaxRange = 50.0;      // racing_env.py:15

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
// Out of line on purpose: IEEE double division and double sincos each expand to a few hundred bytes of SASS,
// and the step kernel has dozens of call sites (ncu: 13 % of stalls were instruction-fetch misses at 130 KB).
#ifndef RK_OUTLINE
#define RK_OUTLINE 0
#endif
#if RK_OUTLINE
#define RK_MAYBE_NOINLINE __noinline__
#else
#define RK_MAYBE_NOINLINE __forceinline__
#endif
__device__ RK_MAYBE_NOINLINE double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ RK_MAYBE_NOINLINE double2 sincos_d(double a) {  // returns (sin, cos) by value: no locals forced to the stack
    double s, c;
    sincos(a, &s, &c);
    return make_double2(s, c);
}
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor_d(v, m));
    return v;
}
// sqrt for culling bounds only (2 instructions; every use adds slack far above its ~2 ulp error)
__device__ __forceinline__ float sqrt_fast(float x) { return x > 0.f ? x * rsqrtf(x) : 0.f; }
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, m));
    return v;
}


// Lexicographic warp minimum of (non-negative double, index) with three REDUX
// instructions: a non-negative IEEE double orders like its bit pattern.
__device__ __forceinline__ int warp_argmin_d(double d, int idx) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_min_sync(kFull, hi);
    const unsigned ml = __reduce_min_sync(kFull, hi == mh ? lo : 0xffffffffu);
    return (int)__reduce_min_sync(kFull, (hi == mh && lo == ml) ? (unsigned)idx : 0xffffffffu);
}

#ifndef RK_L2_UNROLL
#define RK_L2_UNROLL 1
#endif
constexpr int kLevel2Unroll = RK_L2_UNROLL;
// chunk work items per warp: StepParams::list_cap entries (set_tracks sizes it for the largest track of the pool, <= 512;
// every KB of shared memory per CTA that is not needed stays L1: 7 CTAs x 26 KB sit above the 164 KB carve-out, 7 x 22.5 KB below)
constexpr int kRing = 64;      // candidate ring of the sweep: < 32 pending before a push of <= 32, evaluated 32 at a time

// Per-warp shared memory: the pose of the warp's 32 cars (structure of arrays,
// one column per lane: conflict free) and the scratch of the culled queries.
struct CarS {
    double x[32], y[32], c[32], s[32], vx[32], vy[32];
    double cx[4][32], cy[4][32];  // corners FL, FR, RR, RL (car.py:31-36)
};
struct CullView {
    float* rows;                  // the slot area seen as floats: car 0's rays of every environment (zero-copy host rows)
    unsigned long long* ray_key;  // [A*R] (fp32 t bits << 32 | segment) of the best wall candidate
    float2* dir32;                // [A*R]
    unsigned short* list;         // [list_cap]
    unsigned* ring;               // [kRing] pending (segment, ray span) candidates of the sweep's level 2
};
// store of a per-environment result; mirrored into the caller's pinned host arena when zero-copy is on (consecutive
// environments are consecutive lanes, so a warp's stores form one PCIe-friendly run)
template <typename T>
__device__ __forceinline__ void out_store(const StepParams& p, T* ptr, size_t i, T v) {
    ptr[i] = v;
    if (p.arena_delta != 0) {
        char* q = reinterpret_cast<char*>(ptr + i);
        if (q >= p.arena_lo && q < p.arena_hi) *reinterpret_cast<T*>(q + p.arena_delta) = v;
    }
}
// (the per-slot area behind CarS holds the culled mode's fp32 ray directions and candidate keys of ONE environment (16 B
//  per (car, ray) slot) and, once the sweeps are done, car 0's rays of up to 32 / A environments for the zero-copy host rows)
__host__ __device__ inline size_t slot_area_bytes(int A, int R) {
    const size_t per_slot = (size_t)A * R * (8 + 8), rows = (size_t)(32 / A) * R * 4;
    return ((per_slot > rows ? per_slot : rows) + 15) / 16 * 16;
}
// behind the chunk list: the winning segment id of each of the warp's 32 cars' R rays (culled mode)
__host__ __device__ inline size_t warp_smem_bytes(int A, int R, int list_cap) {
    // (the candidate ring exists for multi envs only: one more KB per CTA would push the single env's 7 CTAs per SM past
    //  the 196 KB shared-memory carve-out and halve its L1 -- measured 0.1705 -> 0.179 ms)
    return (sizeof(CarS) + slot_area_bytes(A, R) + (size_t)list_cap * 2 + (size_t)32 * R * 2 + (A > 1 ? kRing * 4 : 0) + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------
// Exact queries (RK_QUERY_EXACT_F64): float64 brute force over the whole table.
// ---------------------------------------------------------------------------
// Waypoint argmin for NQ points (track.py:150-152): lanes stride over the
// waypoints; ties resolve to the lowest index, as numpy's argmin does.
template <int NQ>
__device__ __forceinline__ void argmin_exact(const TrackPool& tp, const TrackMeta& tm, const double* qx,
                                             const double* qy, int lane, int* out_idx) {
    double best[NQ];
    int bi[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { best[q] = INFINITY; bi[q] = 0x7fffffff; }
    const double* wx = tp.wx + tm.wp_off;
    const double* wy = tp.wy + tm.wp_off;
    for (int i = lane; i < tm.n_wp; i += 32) {
        const double px = wx[i], py = wy[i];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const double dx = dsub(px, qx[q]), dy = dsub(py, qy[q]);
            const double d = dadd(dmul(dx, dx), dmul(dy, dy));
            if (d < best[q]) { best[q] = d; bi[q] = i; }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) out_idx[q] = warp_argmin_d(best[q], bi[q]);
}

// One ray against one segment, the reference's test (track.py:176-195 /
// multi_track.py:28-44): hit iff |dotp| large enough, t = cross/dotp >= 0 and
// 0 <= s = dv/dotp <= 1.  Returns t if hit, +inf otherwise.
// FAST: the returned distance (an observation, tolerance 1e-6 on the normalised reading) comes from a Newton
// reciprocal (<= 2 ulp) instead of the IEEE division routine; WHETHER the segment is hit is decided exactly either way.
__device__ __forceinline__ double ddiv_fast(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));   // ~20 good bits
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
template <bool FAST = false>
__device__ __forceinline__ double ray_segment(double v1x, double v1y, double v2x, double v2y, double cross,
                                              double v3x, double v3y, double min_abs_dot) {
    const double dotp = dadd(dmul(v2x, v3x), dmul(v2y, v3y));
    const double dv = dadd(dmul(v1x, v3x), dmul(v1y, v3y));
    const double adot = fabs(dotp), adv = fabs(dv);
    if (adot >= min_abs_dot && adv <= adot * (1.0 + 1e-12)) {
        // t = cross/dotp >= 0 and s = dv/dotp >= 0 are sign statements about correctly
        // rounded quotients; s <= 1 follows from |dv| <= |dotp| (rounding is monotonic) and
        // needs the actual quotient only in the one-ulp band above it.
        const bool neg = dotp < 0.0;
        const bool t_ok = (cross == 0.0) || ((cross < 0.0) == neg);
        const bool s_ok = (dv == 0.0) || ((dv < 0.0) == neg);
        if (t_ok && s_ok && (adv <= adot || ddiv(dv, dotp) <= 1.0)) return FAST ? ddiv_fast(cross, dotp) : ddiv(cross, dotp);
    }
    return INFINITY;
}
// track.py:182 keeps |dotp| > 1e-10; multi_track.py:35 drops |dotp| < 1e-10
constexpr double kWallMinDot = 1.0000000000000002e-10;  // the double after 1e-10: ">" written as ">="
constexpr double kEdgeMinDot = 1e-10;

constexpr int kRayBlock = 8;  // rays whose running minima are kept in registers at once

// nr <= kRayBlock rays of one origin against all walls of a track, exact.
__device__ __forceinline__ void raycast_walls_exact(const TrackPool& tp, const TrackMeta& tm, double ox, double oy,
                                                    const double* v3x, const double* v3y, int nr, int lane,
                                                    double* best) {
    const double* sx = tp.sx + 2 * (size_t)tm.wp_off;
    const double* sy = tp.sy + 2 * (size_t)tm.wp_off;
    const double* v2x = tp.v2x + 2 * (size_t)tm.wp_off;
    const double* v2y = tp.v2y + 2 * (size_t)tm.wp_off;
    const int S = 2 * tm.n_wp;
    for (int i = lane; i < S; i += 32) {
        const double ax = v2x[i], ay = v2y[i];
        const double v1x = dsub(ox, sx[i]), v1y = dsub(oy, sy[i]);
        const double cross = dsub(dmul(ax, v1y), dmul(ay, v1x));
#pragma unroll
        for (int k = 0; k < kRayBlock; ++k)
            if (k < nr) best[k] = fmin(best[k], ray_segment(v1x, v1y, ax, ay, cross, v3x[k], v3y[k], kWallMinDot));
    }
}

// One ray against all walls, exact: the rare re-scan of the culled path.
__device__ RK_MAYBE_NOINLINE double raycast_wall_exact_one(const TrackPool& tp, const TrackMeta& tm, double ox, double oy,
                                                      double v3x, double v3y, int lane) {
    const double* sx = tp.sx + 2 * (size_t)tm.wp_off;
    const double* sy = tp.sy + 2 * (size_t)tm.wp_off;
    const double* v2x = tp.v2x + 2 * (size_t)tm.wp_off;
    const double* v2y = tp.v2y + 2 * (size_t)tm.wp_off;
    double best = INFINITY;
    for (int i = lane; i < 2 * tm.n_wp; i += 32) {
        const double ax = v2x[i], ay = v2y[i];
        const double v1x = dsub(ox, sx[i]), v1y = dsub(oy, sy[i]);
        best = fmin(best, ray_segment(v1x, v1y, ax, ay, dsub(dmul(ax, v1y), dmul(ay, v1x)), v3x, v3y, kWallMinDot));
    }
    return warp_min_d(best);
}

// The four edges of every other car (multi_track.py:10-24), exact, for ONE ray.
template <bool FAST = false>
__device__ __forceinline__ double raycast_car_edges(const CarS& S, int base, int A, double ox, double oy,
                                                    double v3x, double v3y) {
    double t = INFINITY;
    for (int oc = 0; oc < A; ++oc) {
        const int l = base + oc;
        const double ddx = dsub(S.x[l], ox), ddy = dsub(S.y[l], oy);
        if (sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy))) < 0.5) continue;  // multi_track.py:13 (the car itself)
        // all four edges lie within sqrt(5) of the car's centre: a ray whose line passes farther from the
        // centre, or that points away from it, can not hit any of them (no effect on results)
        if (fabs(ddx * v3x + ddy * v3y) > 2.23606797750 + 1e-6 || ddx * v3y - ddy * v3x < -(2.23606797750 + 1e-6)) continue;
#pragma unroll 1
        for (int ed = 0; ed < 4; ++ed) {
            const double ex0 = S.cx[ed][l], ey0 = S.cy[ed][l];
            const double ax = dsub(S.cx[(ed + 1) & 3][l], ex0), ay = dsub(S.cy[(ed + 1) & 3][l], ey0);
            const double v1x = dsub(ox, ex0), v1y = dsub(oy, ey0);
            t = fmin(t, ray_segment<FAST>(v1x, v1y, ax, ay, dsub(dmul(ax, v1y), dmul(ay, v1x)), v3x, v3y, kEdgeMinDot));
        }
    }
    return t;
}

// ---------------------------------------------------------------------------
// Culled queries (RK_QUERY_CULLED).  Tables are fp32 relative to the track's
// bbox centre; kChunk consecutive waypoints / boundary segments (arc-length
// ordered, hence spatially compact) share a bounding circle.  Culling never
// removes a possible winner: every bound carries explicit slack for the fp32
// rounding of tables, query and arithmetic, and winners are re-evaluated in
// float64 (with an exact re-scan if float64 rejects the fp32 candidate).
// ---------------------------------------------------------------------------
constexpr float kHalfDiag = 2.2361f;   // >= sqrt(2^2 + 1^2): corner distance from the car centre
constexpr float kPerpSlack = 3e-4f;    // bound on the fp32 error of a point-to-ray-line distance (|p| <~ 250)
constexpr float kFrontSlack = 1e-3f;
constexpr unsigned long long kNoKey = ~0ull;

// Waypoint argmin for the car centre + 4 corners, float64 scan: one bounding-circle pass picks
// the chunks that can hold the nearest waypoint of ANY of the five points, then
// those chunks are scanned exactly in float64 (half a warp per chunk).  This is the slow path of
// argmin_culled5 / argmin_lane5 below (taken when the fp32 search leaves more than one candidate for a point).
struct CarS;
__device__ __noinline__ void argmin_culled5_f64(const TrackPool& tp, const TrackMeta& tm, const CarS& S, int l,
                                                int lane, unsigned short* list, int* out_idx);
__device__ __forceinline__ void argmin_culled5_f64_body(const TrackPool& tp, const TrackMeta& tm, const double* qx,
                                                const double* qy, int lane, unsigned short* list, int* out_idx) {
    const float4* wch = tp.wchunk + tm.wchunk_off;
    const float cx0 = (float)(qx[0] - tm.org_x), cy0 = (float)(qy[0] - tm.org_y);
    const int nwc = tm.n_wchunk;
    float U = INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < nwc; c0 += 32) {
        const int ci = c0 + lane;
        if (ci < nwc) {
            const float4 cc = wch[ci];
            const float dx = cc.x - cx0, dy = cc.y - cy0;
            U = fminf(U, sqrt_fast(dx * dx + dy * dy) + cc.z);
        }
    }
    // the nearest waypoint of the centre is within U; every corner is within kHalfDiag of
    // the centre, so a chunk whose closest possible waypoint is farther than
    // U + 2*kHalfDiag from the centre can not hold the argmin of any of the five points
    const float thr = warp_min_f(U) + 2.f * kHalfDiag + 2e-2f;
    int count = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
    for (int c0 = 0; c0 < nwc; c0 += 32) {
        const int ci = c0 + lane;
        bool keep = false;
        if (ci < nwc) {
            const float4 cc = wch[ci];
            const float dx = cc.x - cx0, dy = cc.y - cy0;
            keep = sqrt_fast(dx * dx + dy * dy) - cc.z <= thr;
        }
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) list[count + __popc(m & lt)] = (unsigned short)ci;
        count += __popc(m);
    }
    __syncwarp();
    double best[5];
    int bi[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { best[q] = INFINITY; bi[q] = 0x7fffffff; }
    const double* wx = tp.wx + tm.wp_off;
    const double* wy = tp.wy + tm.wp_off;
    const int half = lane >> 4, j = lane & 15;
#pragma unroll 1
    for (int it = 0; it < count; it += 2) {
        const int my = it + half;
        if (my < count) {
            const int i = (int)list[my] * kChunk + j;
            if (i < tm.n_wp) {
                const double px = wx[i], py = wy[i];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const double dx = dsub(px, qx[q]), dy = dsub(py, qy[q]);
                    const double d = dadd(dmul(dx, dx), dmul(dy, dy));
                    if (d < best[q] || (d == best[q] && i < bi[q])) { best[q] = d; bi[q] = i; }
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 5; ++q) out_idx[q] = warp_argmin_d(best[q], bi[q]);
}

// atan2 for the angular sweep: Abramowitz & Stegun 4.4.49 (|error| <= 2e-8 on
// [0, 1]) plus octant reconstruction; the result only bins points between rays
// and its error is part of the binning slack.  (A 3-term polynomial with 6e-4 rad of error was measured: five FMAs
// fewer per boundary point, but every ray that passes within that angle of a boundary vertex then gets a false
// candidate, float64 rejects the winner and the warp re-scans the ray exactly -- 0.290 -> 0.349 ms.)
__device__ __forceinline__ float sweep_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    const float a = __fdividef(mn, mx), s = a * a;
    float r = 0.0028662257f;
    r = fmaf(r, s, -0.0161657367f);
    r = fmaf(r, s, 0.0429096138f);
    r = fmaf(r, s, -0.0752896400f);
    r = fmaf(r, s, 0.1065626393f);
    r = fmaf(r, s, -0.1420889944f);
    r = fmaf(r, s, 0.1999355085f);
    r = fmaf(r, s, -0.3333314528f);
    r = fmaf(r * s, a, a);
    if (ay > ax) r = 1.57079632679f - r;
    if (x < 0.f) r = 3.14159265359f - r;
    return copysignf(r, y);
}

// Polar angle for BINNING boundary points between rays (level 2): atan(a) ~ a (pi/4 + 0.273 (1 - a)) on [0, 1], |error| <=
// 3.8e-3 rad (kBinAngleErr, part of the binning slack), two FMAs instead of nine.  A coarse angle only ADDS (segment, ray)
// candidates next to a bin edge; every candidate is verified against its ray (straddle test) before it may post a key, so
// the extra ones cost an evaluation slot each and never become false winners.
constexpr float kBinAngleErr = 4.0e-3f;
__device__ __forceinline__ float sweep_angle_coarse(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    const float a = __fdividef(mn, mx);
    float r = a * fmaf(a, -0.273f, 1.0583982f);
    if (ay > ax) r = 1.57079632679f - r;
    if (x < 0.f) r = 3.14159265359f - r;
    return copysignf(r, y);
}

// fp32 candidate search of the R rays of one car against the walls, as an
// angular sweep.  The rays are uniformly spaced in angle around the heading
// (np.linspace, racing_env.py:45 / multi_racing_env.py:50), so a boundary
// point's polar angle in the car frame, scaled to "ray index" units u, says
// between which rays it lies; a segment (p, q) can only be hit by the rays whose
// index lies in [min(u_p, u_q), max(u_p, u_q)].
//   Level 1: lane <-> boundary chunk: keep the chunks whose bounding circle
//            reaches into the sensor cone.
//   Level 2: half-warp <-> chunk, lane j <-> point j of its 16 points; lane j+1
//            holds the end point of lane j's segment (one shuffle).  Candidate
//            (ray, segment) pairs post (t, segment) to the ray's key.
// Chunks are visited in distance shells (StepParams::shell).  From the second shell on a
// chunk is kept only if one of the rays pointing into its circle still has no
// candidate, or one that lies beyond the chunk's nearest point: a car inside
// the closed corridor has most rays blocked nearby, so far chunks are skipped
// without changing any ray's nearest hit.  The multi env stops at its 50-unit
// clamp (multi_track.py:8,26).
// Segments seen under ~180 degrees or closer than 0.5 (the origin practically on
// the wall) fall back to an explicit straddle test against every ray.

template <int KIND>
__device__ __forceinline__ void raycast_walls_culled(const TrackPool& tp, const TrackMeta* tmp_, const StepParams& p,
                                                     double oxd, double oyd, double hcd, double hsd, int slot0,
                                                     int lane, const CullView& cv) {
    // (only the six fields needed are read: a by-value copy of the 136-byte record costs 30 loads per environment)
    const float4* bch = tp.bchunk + tmp_->bchunk_off;
    const float2* bpt = tp.bpt + tmp_->bpt_off;
    const float ox = (float)(oxd - tmp_->org_x), oy = (float)(oyd - tmp_->org_y);
    const float hc = (float)hcd, hs = (float)hsd;
    const int nb = tmp_->n_bchunk, nrc = nb >> 1, N = tmp_->n_wp, R = p.R;
    const unsigned lt = (1u << lane) - 1u;
    const int half = lane >> 4, j = lane & 15;
    const float2* dirs = cv.dir32 + slot0;
    unsigned long long* keys = cv.ray_key + slot0;
    const float inv_dphi = p.inv_dphi, u_off = p.cone_half * inv_dphi;
    constexpr bool kRingMode = KIND == RK_ENV_MULTI;   // (see level 2)
    const float wrap_thr = (3.1405926f - (kRingMode ? 2.f * kBinAngleErr : 0.f)) * inv_dphi;   // (the ring mode bins coarse angles)
    const float range = (KIND == RK_ENV_MULTI) ? 50.01f : INFINITY;
    for (int pass = 0; pass < p.n_shells; ++pass) {
        // shell of this pass: chunks whose nearest possible point lies in (lo, hi]
        const float lo = (pass == 0) ? -INFINITY : p.shell[pass - 1];
        const float hi = fminf(p.shell[pass], range);
        if (!(hi > lo)) break;
        // ---- level 1 ----
        int count = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < nb; c0 += 32) {
            const int ci = c0 + lane;
            bool keep = false;
            if (ci < nb) {
                const float4 cc = bch[ci];
                const float rx = cc.x - ox, ry = cc.y - oy, rr = cc.z;
                const float lx = rx * hc + ry * hs, ly = ry * hc - rx * hs;  // car frame
                const float dc = sqrt_fast(rx * rx + ry * ry), dmin = dc - rr;   // no point of the chunk is closer than dmin
                keep = (p.cone_cos * fabsf(ly) - p.cone_sin * lx <= rr) &&   // circle reaches into the cone |angle| <= H
                       dmin > lo && dmin <= hi;
                if (keep && pass > 0 && dc > rr) {
                    // per-ray pruning: the circle subtends <= asin(rr/dc) <= (pi/2) rr/dc around its centre;
                    // keep it only if a ray in that fan has no candidate yet or one farther than dmin
                    const float uc = sweep_atan2(ly, lx) * inv_dphi + u_off;
                    const float du = (1.5708f * rr / dc + 1e-3f) * inv_dphi;
                    const int klo = max(0, (int)ceilf(uc - du)), khi = min(R - 1, (int)floorf(uc + du));
                    bool open = false;
                    for (int k = klo; k <= khi; ++k) {
                        const unsigned tb = (unsigned)(keys[k] >> 32);  // 0xffffffff (no candidate) is a NaN pattern: test first
                        open = open || tb == 0xffffffffu || __uint_as_float(tb) * 1.0001f + 1e-2f > dmin;
                    }
                    keep = open;
                }
            }
            const unsigned m = __ballot_sync(kFull, keep);
            if (keep) cv.list[count + __popc(m & lt)] = (unsigned short)ci;
            count += __popc(m);
        }
        __syncwarp();
        // ---- level 2 ----
        // Multi env (one pass over ~24 chunks per car): segments that can be crossed by a ray post ONE item (segment,
        // first ray, ray count) to the warp's ring; whenever 32 items are pending they are evaluated one per lane -- the
        // candidate loop inside the chunk loop runs at ~3 of 32 lanes -- and the binning angle is the coarse one, since
        // every candidate is verified against its ray anyway.  Measured 0.277 -> 0.268 ms at 65,536 two-car envs.  The
        // single env (two short shells per car, whose candidates the second shell's pruning needs at once) keeps the
        // in-loop evaluation and the accurate angle: the ring's per-chunk push and per-shell flush cost it 13 %.
        unsigned head = 0, tail = 0;
        auto evaluate = [&](unsigned n_items) {
            if (lane < n_items) {
                const unsigned it = cv.ring[(head + lane) & (kRing - 1)];
                const int seg = (int)(it & 0xffffu), k0 = (int)((it >> 16) & 0x7fu), kn = (int)((it >> 23) & 0xffu);
                const int side = seg >= N, pt = seg - side * N;
                const float2 P = bpt[side * (N + 1) + pt], Q = bpt[side * (N + 1) + pt + 1];
                const float px = P.x - ox, py = P.y - oy, qx = Q.x - ox, qy = Q.y - oy;
                const float vx = qx - px, vy = qy - py, num = px * vy - py * vx;
                for (int k = k0; k < k0 + kn; ++k) {
                    const float2 d = dirs[k];
                    const float cp = d.x * py - d.y * px, cq = d.x * qy - d.y * qx;  // signed distances to the ray's line
                    const bool straddle = (cp <= kPerpSlack && cq >= -kPerpSlack) || (cp >= -kPerpSlack && cq <= kPerpSlack);
                    const float tp_ = px * d.x + py * d.y, tq_ = qx * d.x + qy * d.y;
                    const float tlo = fminf(tp_, tq_), thi = fmaxf(tp_, tq_);
                    if (!straddle || thi < -kFrontSlack) continue;
                    const float den = d.x * vy - d.y * vx;
                    float t = (fabsf(den) > 1e-12f) ? __fdividef(num, den) : tlo;
                    t = fmaxf(fminf(fmaxf(t, tlo), thi), 0.f);  // the crossing lies between the end points
                    atomicMin(&keys[k], ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)seg);
                }
            }
        };
#pragma unroll kLevel2Unroll
        for (int it = 0; it < count; it += 2) {
            const int my = it + half;
            const bool act = my < count;
            const int ci = act ? (int)cv.list[my] : 0;
            const int side = ci >= nrc;
            const int pt = (ci - side * nrc) * kRaySegs + j;  // point index in the closed row (0..N)
            float px = 1e6f, py = 1e6f;
            if (act && pt <= N) {
                const float2 P = bpt[side * (N + 1) + pt];
                px = P.x - ox; py = P.y - oy;
            }
            const float ly_ = py * hc - px * hs, lx_ = px * hc + py * hs;
            const float u = (kRingMode ? sweep_angle_coarse(ly_, lx_) : sweep_atan2(ly_, lx_)) * inv_dphi + u_off;
            const float qx = __shfl_down_sync(kFull, px, 1), qy = __shfl_down_sync(kFull, py, 1);
            const float un = __shfl_down_sync(kFull, u, 1);
            int klo = 1, khi = 0;
            bool slow = false;
            if (act && j < kRaySegs && pt < N) {
                const float m2 = fminf(px * px + py * py, qx * qx + qy * qy);
                slow = fabsf(un - u) > wrap_thr || m2 < 0.25f;
                // angular error budget: the coarse angle + table/origin rounding (<= 1.6e-5 / distance) + heading rounding
                const float slack = ((kRingMode ? kBinAngleErr : 0.f) + 3e-6f + 1.6e-5f * rsqrtf(m2)) * inv_dphi;
                klo = slow ? 0 : max(0, (int)ceilf(fminf(u, un) - slack));
                khi = slow ? R - 1 : min(R - 1, (int)floorf(fmaxf(u, un) + slack));
            }
            if (!kRingMode) {
                for (int k = klo; k <= khi; ++k) {
                    const float2 d = dirs[k];
                    const float tp_ = px * d.x + py * d.y, tq_ = qx * d.x + qy * d.y;
                    const float tlo = fminf(tp_, tq_), thi = fmaxf(tp_, tq_);
                    if (slow) {
                        const float cp = d.x * py - d.y * px, cq = d.x * qy - d.y * qx;  // signed distances to the ray's line
                        const bool straddle = (cp <= kPerpSlack && cq >= -kPerpSlack) ||
                                              (cp >= -kPerpSlack && cq <= kPerpSlack);
                        if (!straddle || thi < -kFrontSlack) continue;
                    }
                    const float vx = qx - px, vy = qy - py;
                    const float den = d.x * vy - d.y * vx;
                    float t = (fabsf(den) > 1e-12f) ? __fdividef(px * vy - py * vx, den) : tlo;
                    t = fmaxf(fminf(fmaxf(t, tlo), thi), 0.f);  // the crossing lies between the end points
                    atomicMin(&keys[k], ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)(side * N + pt));
                }
                continue;
            }
            const bool has = klo <= khi;
            const unsigned m = __ballot_sync(kFull, has);
            if (m) {
                if (has)
                    cv.ring[(tail + __popc(m & lt)) & (kRing - 1)] =
                        (unsigned)(side * N + pt) | ((unsigned)klo << 16) | ((unsigned)(khi - klo + 1) << 23);
                tail += __popc(m);
                if (tail - head >= 32u) {
                    __syncwarp();
                    evaluate(32u);
                    head += 32u;
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        if (kRingMode && tail != head) evaluate(tail - head);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Grid queries (RK_QUERY_GRID): every lane works on its own car / its own ray.
// ---------------------------------------------------------------------------
// Waypoint argmin of the car centre + 4 corners (track.py:150-152, five times per car.py:79-80 / track.py:163-171),
// ONE CAR PER LANE.  The result must be numpy's float64 argmin, first index on ties, because it decides progress,
// checkpoints and the wall test.  The search itself runs in fp32:
//   1. the distance from the centre to ANY waypoint bounds the nearest distance from above; the waypoint found one
//      step earlier (`hint`) is within a few indices of the new one, so the bound is tight.  The lane walks the
//      bounding circles and remembers up to six chunks that can hold the nearest waypoint of the centre or of a
//      corner (corners lie within kHalfDiag of the centre);
//   2. fp32 squared distances of the five points to those chunks' waypoints, with best and second-best per point;
//   3. a point whose fp32 minimum is separated from every other waypoint by more than the rounding bound has exactly
//      one candidate, and that candidate IS the float64 argmin -- no float64 arithmetic needed.  Otherwise (near
//      ties: the point is within ~1e-4 of the bisector of two waypoints; more than six chunks) the function returns
//      false and the cooperative float64 scan (argmin_culled5_f64) decides.
// Rounding bound: table and query are rounded to fp32 relative to the bbox centre (|coordinate| <= diag/2, so each is
// off by at most diag * 2^-25), hence |d32 - d64| <= 4 sqrt(2) diag 2^-25 d + 4e-7 d^2 for a squared distance; two
// such errors meet in a comparison.  The bound used below is more than twice that.
__device__ __forceinline__ bool argmin_lane5(const TrackPool& tp, const TrackMeta* tm, double x, double y,
                                             const double* cxs, const double* cys, int hint, int* out_idx) {
    const float4* wch = tp.wchunk + tm->wchunk_off;
    const float2* wpt = tp.wpt + tm->wp_off;
    const int n_wp = tm->n_wp, nwc = tm->n_wchunk;
    const double orgx = tm->org_x, orgy = tm->org_y;
    float fx[5], fy[5];
    fx[0] = (float)(x - orgx); fy[0] = (float)(y - orgy);
#pragma unroll
    for (int q = 0; q < 4; ++q) { fx[q + 1] = (float)(cxs[q] - orgx); fy[q + 1] = (float)(cys[q] - orgy); }
    const float2 wh = wpt[min(max(hint, 0), n_wp - 1)];
    const float hx = wh.x - fx[0], hy = wh.y - fy[0];
    const float thr = sqrt_fast(hx * hx + hy * hy) + 2.f * kHalfDiag + 2e-2f;
    unsigned long long keep = 0ull;   // up to six 10-bit chunk ids
    int nkeep = 0;
#pragma unroll 1
    for (int c = 0; c < nwc; ++c) {
        const float4 cc = wch[c];
        const float dx = cc.x - fx[0], dy = cc.y - fy[0];
        const float lim = thr + cc.z;
        if (dx * dx + dy * dy <= lim * lim) {
            if (nkeep < 6) keep |= (unsigned long long)c << (10 * nkeep);
            ++nkeep;
        }
    }
    float best[5], second[5];
    int bi[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { best[q] = INFINITY; second[q] = INFINITY; bi[q] = 0; }
    const int nscan = min(nkeep, 6);
#pragma unroll 1
    for (int k = 0; k < nscan; ++k) {
        const int i0 = (int)((keep >> (10 * k)) & 1023ull) * kChunk;
        const int i1 = min(i0 + kChunk, n_wp);
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const float2 P = wpt[i];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float dx = P.x - fx[q], dy = P.y - fy[q];
                const float d = fmaf(dx, dx, dy * dy);
                second[q] = fminf(second[q], fmaxf(d, best[q]));
                if (d < best[q]) { best[q] = d; bi[q] = i; }
            }
        }
    }
    const float c_err = (float)tm->max_track_distance * 6e-7f;
    bool ok = nkeep >= 1 && nkeep <= 6;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const float m = best[q];
        ok = ok && second[q] > m + c_err * sqrt_fast(m) + 2e-6f * m + 1e-9f;
        out_idx[q] = bi[q];
    }
    return ok;
}

// One ray through the track's uniform grid (Amanatides & Woo traversal), ONE RAY PER LANE: returns the two nearest
// fp32 hits (distance, segment id) among the boundary segments listed in the cells the ray crosses.  The hit test is
// the reference's (track.py:176-195) with slack: t >= -eps, -eps <= s <= 1 + eps, so rounding can only add
// candidates; the caller re-evaluates them in float64.  The walk ends once the best hit lies inside the cells
// already visited, at `tmax`, or at the grid's edge.  inside = false: the origin is outside the grid.
struct RayHit { float t0, t1; int s0, s1; bool inside; };
__device__ __forceinline__ RayHit grid_ray(const TrackPool& tp, const TrackMeta* tm, float ox, float oy, float dx,
                                           float dy, float tmax) {
    RayHit h;
    h.t0 = INFINITY; h.t1 = INFINITY; h.s0 = -1; h.s1 = -1;
    const float cell = tm->gcell, inv = tm->ginv, gx0 = tm->gx0, gy0 = tm->gy0;
    const int nx = tm->gnx, ny = tm->gny;
    int ix = (int)floorf((ox - gx0) * inv), iy = (int)floorf((oy - gy0) * inv);
    h.inside = (unsigned)ix < (unsigned)nx && (unsigned)iy < (unsigned)ny;
    if (!h.inside) return h;
    const uint32_t* cells = tp.gcell + tm->gcell_off;
    const uint16_t* lst = tp.glist + tm->glist_off;
    const float4* segs = tp.bseg + 2 * (size_t)tm->wp_off;
    const int stx = dx >= 0.f ? 1 : -1, sty = dy >= 0.f ? 1 : -1;
    const float rdx = fabsf(dx) > 1e-20f ? 1.f / dx : copysignf(1e20f, dx), rdy = fabsf(dy) > 1e-20f ? 1.f / dy : copysignf(1e20f, dy);
    float tmx = (gx0 + (float)(ix + (stx > 0)) * cell - ox) * rdx;   // ray parameter at the next x / y cell boundary
    float tmy = (gy0 + (float)(iy + (sty > 0)) * cell - oy) * rdy;
    const float tdx = cell * fabsf(rdx), tdy = cell * fabsf(rdy);
    // Outer iteration = one NON-EMPTY cell: a tight loop first walks over the empty cells in between (the inside of
    // the corridor holds no segments), a second tight loop tests the cell's segments.  Lanes of a warp run rays of
    // similar length of different cars (see the ray order below), so both inner loops have similar trip counts.
    bool stop = false;
    uint32_t c = cells[iy * nx + ix];
    float t_exit = fminf(tmx, tmy);
#pragma unroll 1
    while (!stop) {
#pragma unroll 1
        while ((c & 1023u) == 0u) {
            const bool xstep = tmx < tmy;   // branch-free step
            ix += xstep ? stx : 0;  iy += xstep ? 0 : sty;
            tmx += xstep ? tdx : 0.f;  tmy += xstep ? 0.f : tdy;
            if (t_exit > tmax || (unsigned)ix >= (unsigned)nx || (unsigned)iy >= (unsigned)ny) { stop = true; break; }
            c = cells[iy * nx + ix];
            t_exit = fminf(tmx, tmy);
        }
        if (stop) break;
#pragma unroll 1
        for (uint32_t k = c >> 10, e = (c >> 10) + (c & 1023u); k < e; ++k) {
            const int sg = lst[k];
            const float4 S = segs[sg];
            const float wx = S.x - ox, wy = S.y - oy;
            const float den = dx * S.w - dy * S.z;        // d x v   (= the reference's dotp)
            const float aden = fabsf(den), sgn = copysignf(1.f, den);
            const float tn = (wx * S.w - wy * S.z) * sgn;  // (w x v) sign(den): t = tn / |den|
            const float sn = (wx * dy - wy * dx) * sgn;    // (w x d) sign(den): s = sn / |den|
            // slack: |w| <~ 300, products carry <~ 1e-4 of absolute error
            if (aden > 1e-12f && tn >= -2e-4f && sn >= -2e-4f && sn <= aden + 2e-4f) {
                const float t = fmaxf(__fdividef(tn, aden), 0.f);
                if (t < h.t0) { h.t1 = h.t0; h.s1 = h.s0; h.t0 = t; h.s0 = sg; }
                else if (t < h.t1 && sg != h.s0) { h.t1 = t; h.s1 = sg; }
            }
        }
        if (h.t0 <= t_exit || t_exit > tmax) break;
        const bool xstep = tmx < tmy;
        ix += xstep ? stx : 0;  iy += xstep ? 0 : sty;
        tmx += xstep ? tdx : 0.f;  tmy += xstep ? 0.f : tdy;
        if ((unsigned)ix >= (unsigned)nx || (unsigned)iy >= (unsigned)ny) break;
        c = cells[iy * nx + ix];
        t_exit = fminf(tmx, tmy);
    }
    return h;
}

// Out-of-line slow path of argmin_culled5 / argmin_lane5: the five query points are re-read from the warp's shared
// pose table and the indices come back through shared memory, so that the caller's registers never spill for its sake.
__device__ __noinline__ void argmin_culled5_f64(const TrackPool& tp, const TrackMeta& tm, const CarS& S, int l,
                                                int lane, unsigned short* list, int* out_idx) {
    double qx[5], qy[5];
    qx[0] = S.x[l]; qy[0] = S.y[l];
#pragma unroll
    for (int k = 0; k < 4; ++k) { qx[k + 1] = S.cx[k][l]; qy[k + 1] = S.cy[k][l]; }
    int idx[5];
    argmin_culled5_f64_body(tp, tm, qx, qy, lane, list, idx);
    if (lane < 5) out_idx[lane] = lane == 0 ? idx[0] : lane == 1 ? idx[1] : lane == 2 ? idx[2] : lane == 3 ? idx[3] : idx[4];
    __syncwarp();
}

// Philox Fisher-Yates over the A car ids of an environment; returns the grid
// slot of car `a` (multi_racing_env.py:127-133).  Every lane of the environment
// draws the same numbers, so no communication is needed.
__device__ __forceinline__ int philox_start_slot(uint64_t seed, int e, uint32_t reset_count, int A, int a) {
    uint32_t perm = 0x76543210u;  // nibble k = car id at grid position k
    for (int i = A - 1; i > 0; --i) {
        uint32_t ctr[4] = {(uint32_t)e, reset_count, (uint32_t)i, 0x736c6f74u};
        philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
        const int j = (int)(((uint64_t)ctr[0] * (uint64_t)(i + 1)) >> 32);
        const uint32_t vi = (perm >> (4 * i)) & 15u, vj = (perm >> (4 * j)) & 15u;
        perm &= ~((15u << (4 * i)) | (15u << (4 * j)));
        perm |= (vj << (4 * i)) | (vi << (4 * j));
    }
    int slot = 0;
    for (int k = 0; k < A; ++k)
        if (((perm >> (4 * k)) & 15u) == (uint32_t)a) slot = k;
    return slot;
}

// ---------------------------------------------------------------------------
#ifndef RK_STEP_MIN_BLOCKS
#define RK_STEP_MIN_BLOCKS 7
#endif
// STAGED: the CTA's environments all run on ONE track (the host groups them, see
// rk_api.cu), whose fp32 search tables -- closed boundary rows, ray-chunk and
// waypoint-chunk circles -- are brought into shared memory with bulk asynchronous
// copies (TMA, cp.async.bulk -> mbarrier) while the warps integrate the dynamics.
template <int KIND, int QUERY, bool STAGED>
__global__ void __launch_bounds__(kWarpsPerCta * 32, RK_STEP_MIN_BLOCKS) step_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long stage_bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A, R = p.R, D = p.D;
    const int epw = p.epw;                                        // environments per warp (<= 32 / A)
    const int gwarp = blockIdx.x * kWarpsPerCta + warp;           // this warp's group of environments
    TrackPool tp = p.trk;
    TrackMeta stm;                                                // STAGED: the CTA's track, offsets re-based to shared memory
    if (STAGED) {
        stm = tp.meta[p.cta_track[blockIdx.x]];
        unsigned char* stage = smem_raw + (size_t)kWarpsPerCta * warp_smem_bytes(A, R, p.list_cap);
        const unsigned b_bpt = 16u * (unsigned)(stm.n_wp + 1), b_bch = 16u * (unsigned)stm.n_bchunk,
                       b_wch = 16u * (unsigned)stm.n_wchunk;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_bpt + b_bch + b_wch) : "memory");
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stage);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(tp.bpt + stm.bpt_off), "r"(b_bpt), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + b_bpt), "l"(tp.bchunk + stm.bchunk_off), "r"(b_bch), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + b_bpt + b_bch), "l"(tp.wchunk + stm.wchunk_off), "r"(b_wch), "r"(bar) : "memory");
        }
        tp.bpt = reinterpret_cast<const float2*>(stage);
        tp.bchunk = reinterpret_cast<const float4*>(stage + b_bpt);
        tp.wchunk = reinterpret_cast<const float4*>(stage + b_bpt + b_bch);
        stm.bpt_off = 0; stm.bchunk_off = 0; stm.wchunk_off = 0;
        __syncthreads();  // the barrier is initialised before any warp polls it (every thread reaches this point)
    }
    int e_base = p.env_begin + gwarp * epw, n_env;
    if (STAGED) {
        n_env = p.group_count[gwarp];
        if (n_env == 0) return;
    } else {
        if (e_base >= p.env_end) return;
        n_env = min(epw, p.env_end - e_base);
    }
    const int* genv = STAGED ? p.group_env + (size_t)gwarp * epw : nullptr;
    unsigned char* wbase = smem_raw + (size_t)warp * warp_smem_bytes(A, R, p.list_cap);
    CarS& S = *reinterpret_cast<CarS*>(wbase);
    CullView cv;
    static_assert(sizeof(CarS) % 16 == 0, "the slot area must stay 16-byte aligned");
    cv.rows = reinterpret_cast<float*>(wbase + sizeof(CarS));
    cv.ray_key = reinterpret_cast<unsigned long long*>(wbase + sizeof(CarS));
    cv.dir32 = reinterpret_cast<float2*>(cv.ray_key + A * R);
    cv.list = reinterpret_cast<unsigned short*>(wbase + sizeof(CarS) + slot_area_bytes(A, R));
    cv.ring = reinterpret_cast<unsigned*>(wbase + sizeof(CarS) + slot_area_bytes(A, R) + (size_t)p.list_cap * 2 + (size_t)32 * R * 2);

    // ---- lane = one car -------------------------------------------------------
    const int g = lane / A, a = lane - g * A;     // environment within the warp, car within the environment
    const bool is_car = g < n_env;
    const int e = STAGED ? genv[is_car ? g : 0] : e_base + (is_car ? g : 0);
    const int base = g * A;                       // first lane of this car's environment
    const bool lead = is_car && a == 0;           // writes the per-environment outputs
    const int c = e * A + a;                      // state index
    const bool agent_major = p.io.layout == RK_LAYOUT_AGENT_MAJOR;
    const size_t ci = agent_major ? (size_t)a * p.E + e : (size_t)c;  // index in the caller's per-car arrays
    const int tid = tp.env_to_track[e];
    const TrackMeta* tmp = tp.meta + tid;  // (scalar fields only: start pose, width, N)
    const int n_wp = tmp->n_wp;
    const double Nd = (double)n_wp;

    // which of {step, reset, observe} applies to this car's environment
    bool stepping = is_car && (p.mode == 0);
    bool resetting = false;
    if (is_car && p.mode == 0 && p.autoreset == RK_AUTORESET_NEXT_STEP && p.st.needs_reset[e]) {
        stepping = false;  // the action is ignored (gymnasium NEXT_STEP)
        resetting = true;
    }
    if (is_car && p.mode == 1) resetting = (p.reset_mask == nullptr) || (p.reset_mask[e] != 0);

    double x = 0, y = 0, ang = 0, vx = 0, vy = 0;
    float last_steer = 0.f;
    int pidx = 0, lpidx = 0, flags = 0, fstep = 0, steps = 0;
    if (is_car) {
        x = p.st.x[c]; y = p.st.y[c]; ang = p.st.ang[c]; vx = p.st.vx[c]; vy = p.st.vy[c];
        last_steer = p.st.last_steer[c];
        pidx = p.st.pidx[c]; lpidx = p.st.lpidx[c]; flags = p.st.flags[c]; fstep = p.st.fstep[c];
        steps = p.st.steps[e];
    }
    double reward = 0.0, delta = 0.0, cs = 1.0, sn = 0.0;
    int placement = 0;
    bool terminated = false, truncated = false;

    if (p.mode == 0) {
        // ---- D: vehicle dynamics (car.py:45-80) ---------------------------------
        const bool moving = stepping && !(flags & F_CRASHED);  // car.py:51-52
        // (zero-copy host face: car 0's pair comes straight from the caller's pinned buffer, 8 coalesced bytes per env,
        //  and is written through so that the device array stays current -- also for envs that ignore it this step)
        float2 av = make_float2(0.f, 0.f);
        if (p.act_host0 != nullptr && is_car && a == 0) {
            av = *reinterpret_cast<const float2*>(p.act_host0 + 2 * (size_t)e);
            *reinterpret_cast<float2*>(const_cast<float*>(p.io.actions) + 2 * ci) = av;
        } else if (stepping) {
            av = *reinterpret_cast<const float2*>(p.io.actions + 2 * ci);
        }
        if (stepping) {
            const float a0 = av.x, a1 = av.y;
            const float steer_f = fminf(fmaxf(a0, -1.f), 1.f);  // racing_env.py:106
            float thr_f;
            if (KIND == RK_ENV_SINGLE)
                thr_f = fminf(fmaxf(a1, 0.f), 1.f);  // racing_env.py:107
            else  // multi_racing_env.py:217, evaluated in float32 for float32 actions
                thr_f = fminf(fmaxf(__fdiv_rn(__fadd_rn(a1, 1.f), 2.f), 0.f), 1.f);
            last_steer = steer_f;
            if (moving) {
                const double steer = (double)steer_f, thr = (double)thr_f;
                double na = dadd(ang, dmul(dmul(steer, 3.0), kDt));  // car.py:54-55
                na = fmod(na, kTwoPi);                                // car.py:56 (python modulo)
                if (na != 0.0) { if (na < 0.0) na = dadd(na, kTwoPi); } else na = 0.0;
                ang = na;
                { const double2 sc = sincos_d(ang); sn = sc.x; cs = sc.y; }
                double vf = dadd(dmul(vx, cs), dmul(vy, sn));            // car.py:59
                double vl = dadd(dmul(vx, -sn), dmul(vy, cs));           // car.py:60
                vf = dmul(dadd(vf, dmul(dmul(thr, 10.0), kDt)), 0.985);  // car.py:61-62
                vl = dmul(dmul(vl, 0.85), 0.9);                          // car.py:63
                vx = dsub(dmul(vf, cs), dmul(vl, sn));                   // car.py:66-67
                vy = dadd(dmul(vf, sn), dmul(vl, cs));
                const double speed = sqrt(dadd(dmul(vx, vx), dmul(vy, vy)));  // car.py:70
                if (speed > kMaxSpeed) {
                    const double scale = ddiv(kMaxSpeed, speed);
                    vx = dmul(vx, scale);
                    vy = dmul(vy, scale);
                }
                x = dadd(x, dmul(vx, kDt));  // car.py:77-78
                y = dadd(y, dmul(vy, kDt));
            }
        }
        if (!moving) { const double2 sc = sincos_d(ang); sn = sc.x; cs = sc.y; }
        // publish pose and corners (car.py:26-43) for the cooperative phases
        S.x[lane] = x; S.y[lane] = y;
        {
            const double lx[4] = {2.0, 2.0, -2.0, -2.0}, ly[4] = {1.0, -1.0, -1.0, 1.0};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                S.cx[k][lane] = dadd(dadd(dmul(cs, lx[k]), dmul(-sn, ly[k])), x);
                S.cy[k][lane] = dadd(dadd(dmul(sn, lx[k]), dmul(cs, ly[k])), y);
            }
        }
        __syncwarp();

        if (STAGED) {  // the staged tables are needed from here on
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar) : "memory");
        }
        // ---- W: closest waypoint of the centre and the 4 corners, one car at a time,
        //         all lanes cooperating (track.py:150-152) ---------------------------
        int cidx1 = 0, cidx2 = 0, cidx3 = 0, cidx4 = 0;
        unsigned todo = __ballot_sync(kFull, moving);
        if (QUERY != RK_QUERY_EXACT_F64 && p.lane_argmin) {
            // every moving car searches on its own lane; the few that end in a near tie go through the cooperative
            // float64 scan below, one at a time
            bool solved = true;
            if (moving) {
                const double cxs[4] = {S.cx[0][lane], S.cx[1][lane], S.cx[2][lane], S.cx[3][lane]};
                const double cys[4] = {S.cy[0][lane], S.cy[1][lane], S.cy[2][lane], S.cy[3][lane]};
                int idx[5];
                solved = argmin_lane5(tp, STAGED ? &stm : tmp, x, y, cxs, cys, lpidx, idx);   // (staged: table offsets are re-based)
                pidx = idx[0]; cidx1 = idx[1]; cidx2 = idx[2]; cidx3 = idx[3]; cidx4 = idx[4];
            }
            todo = __ballot_sync(kFull, moving && !solved);
        }
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const TrackMeta tm = STAGED ? stm : tp.meta[__shfl_sync(kFull, tid, l)];
            double qx[5], qy[5];
            qx[0] = S.x[l]; qy[0] = S.y[l];
#pragma unroll
            for (int k = 0; k < 4; ++k) { qx[k + 1] = S.cx[k][l]; qy[k + 1] = S.cy[k][l]; }
            int idx[5];
            if (QUERY != RK_QUERY_EXACT_F64) {
                int* res = reinterpret_cast<int*>(cv.list + p.list_cap - 16);   // past any chunk list (list_cap = largest list + 16, rounded up)
                argmin_culled5_f64(tp, tm, S, l, lane, cv.list, res);
#pragma unroll
                for (int q = 0; q < 5; ++q) idx[q] = res[q];
                __syncwarp();
            } else
                argmin_exact<5>(tp, tm, qx, qy, lane, idx);
            if (lane == l) { pidx = idx[0]; cidx1 = idx[1]; cidx2 = idx[2]; cidx3 = idx[3]; cidx4 = idx[4]; }  // car.py:79
        }
        // ---- C: wall test of the 4 corners (track.py:163-171), every car in parallel
        if (moving) {
            const int off = tmp->wp_off;
            const double width = tmp->width;
            const int cidx[4] = {cidx1, cidx2, cidx3, cidx4};
            bool crashed = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = off + cidx[k];
                const double dist = fabs(dadd(dmul(dsub(S.cx[k][lane], tp.wx[i]), tp.nx[i]),
                                              dmul(dsub(S.cy[k][lane], tp.wy[i]), tp.ny[i])));
                crashed |= dist > width;
            }
            if (crashed) flags |= F_CRASHED;
        }

        // ---- X: car-car collisions (multi_racing_env.py:222-231): every car tests
        //         itself against every other car of its environment ------------------
        double touching = 0.0;
        if (KIND == RK_ENV_MULTI && A > 1 && stepping) {
            for (int o = 0; o < A; ++o) {
                if (o == a) continue;
                const int li = base + min(a, o), lj = base + max(a, o);  // the reference's (i, j), i < j
                bool hit = true;  // multi_car.py:16-43: 4 axes, strict separation test
                {   // both rectangles lie within sqrt(5) of their centres: farther apart than 2 sqrt(5) they cannot touch
                    const double ddx = S.x[lj] - S.x[li], ddy = S.y[lj] - S.y[li];
                    if (ddx * ddx + ddy * ddy > 20.001) hit = false;
                }
#pragma unroll 1
                for (int ax = 0; hit && ax < 4; ++ax) {
                    const int lo = (ax < 2) ? li : lj, k0 = ax & 1;
                    const double ex = dsub(S.cx[k0 + 1][lo], S.cx[k0][lo]), ey = dsub(S.cy[k0 + 1][lo], S.cy[k0][lo]);
                    const double nx = -ey, ny = ex;
                    double amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double pa = dadd(dmul(S.cx[k][li], nx), dmul(S.cy[k][li], ny));
                        const double pb = dadd(dmul(S.cx[k][lj], nx), dmul(S.cy[k][lj], ny));
                        amin = fmin(amin, pa); amax = fmax(amax, pa);
                        bmin = fmin(bmin, pb); bmax = fmax(bmax, pb);
                    }
                    if (amax < bmin || bmax < amin) hit = false;
                }
                if (hit) {
                    vx = dmul(vx, 0.92);
                    vy = dmul(vy, 0.92);
                    touching = dadd(touching, -5.0);
                }
            }
        }
        if (stepping) steps += 1;  // racing_env.py:110 / multi_racing_env.py:233

        // ---- reward and episode logic ------------------------------------------------
        if (stepping) {
            const double prog = ddiv((double)pidx, Nd), lprog = ddiv((double)lpidx, Nd);  // track.py:161
            delta = dsub(prog, lprog);  // racing_env.py:112-116 / multi:159-163
            if (lprog > 0.9 && prog < 0.1) delta = dadd(dsub(1.0, lprog), prog);
            else if (lprog < 0.1 && prog > 0.9) delta = -dadd(dsub(1.0, prog), lprog);
            const bool crashed = flags & F_CRASHED;
            const double speed = sqrt(dadd(dmul(vx, vx), dmul(vy, vy)));
            const double ratio = clipd(ddiv(speed, kMaxSpeed), 0.0, 1.0);
            reward = dmul(delta, 200.0);
            const double cp_bonus = (KIND == RK_ENV_SINGLE) ? 20.0 : 25.0;
            if (KIND == RK_ENV_MULTI && !crashed && delta > 0.0) reward = dadd(reward, dmul(ratio, 18.0));  // multi:169-172
            if (!(flags & F_CP25) && 0.25 <= prog && prog < 0.35) { flags |= F_CP25; reward = dadd(reward, cp_bonus); }
            if ((flags & F_CP25) && !(flags & F_CP50) && 0.50 <= prog && prog < 0.60) { flags |= F_CP50; reward = dadd(reward, cp_bonus); }
            if ((flags & F_CP50) && !(flags & F_CP75) && 0.75 <= prog && prog < 0.85) { flags |= F_CP75; reward = dadd(reward, cp_bonus); }
            const bool all_cp = (flags & (F_CP25 | F_CP50 | F_CP75)) == (F_CP25 | F_CP50 | F_CP75);
            const bool fin = all_cp && lprog > 0.9 && prog < 0.1 && delta > 0.0;
            if (KIND == RK_ENV_SINGLE) {
                if (!crashed && delta > 0.0) reward = dadd(reward, dmul(ratio, p.speed_weight));  // racing_env.py:137-140
                if (crashed) reward = dsub(reward, 60.0);                                        // :142-143
                if (fin) {                                                                       // :145-150
                    flags |= F_FINISHED;
                    reward = dadd(reward, 100.0);
                    reward = dadd(reward, fmax(0.0, dsub(200.0, ddiv((double)steps, 10.0))));
                }
            } else {
                if (fin) {  // multi:186-190
                    flags |= F_FINISHED;
                    fstep = steps;
                    reward = dadd(reward, dadd(100.0, fmax(0.0, dsub(300.0, ddiv((double)steps, 15.0)))));
                }
                if (crashed && !(flags & F_HAS_CRASHED)) {  // multi:192-194
                    reward = dsub(reward, 160.0);
                    flags |= F_HAS_CRASHED;
                }
                reward = dadd(reward, touching);  // multi:240
            }
        }
        // environment-level termination from the flags of its A cars
        const unsigned grp_mask = ((A >= 32) ? kFull : ((1u << A) - 1u));
        const unsigned fin_b = (__ballot_sync(kFull, stepping && (flags & F_FINISHED)) >> base) & grp_mask;
        const unsigned crash_b = (__ballot_sync(kFull, stepping && (flags & F_CRASHED)) >> base) & grp_mask;
        if (stepping) {
            if (KIND == RK_ENV_SINGLE)
                terminated = (fin_b | crash_b) != 0u;  // racing_env.py:161
            else
                terminated = (fin_b != 0u) || (crash_b == grp_mask);  // multi:247-249
            truncated = steps >= p.max_steps;
        }
        if (KIND == RK_ENV_MULTI) {
            // place() multi:198-211: descending (score, idx); exact ties go to the higher index
            double score = 0.0;
            if (stepping) {
                const double prog = ddiv((double)pidx, Nd);
                score = dadd(dadd(dadd((flags & F_FINISHED) ? 10000.0 : 0.0, dmul(prog, 100.0)),
                                  (flags & F_CRASHED) ? 0.0 : 10.0),
                             ddiv(1.0, (double)(fstep ? fstep : 10000)));
            }
            int ahead = 0;
            for (int o = 0; o < A; ++o) {
                const double so = __shfl_sync(kFull, score, (base + o) & 31);
                if (o != a && (so > score || (so == score && o > a))) ++ahead;
            }
            if (stepping && (terminated || truncated)) {
                placement = ahead + 1;
                if (placement == 1) reward = dadd(reward, 250.0);  // multi:256-257
            }
        }
        if (stepping) lpidx = pidx;  // racing_env.py:165 / multi:266-267
    }

    // ---- episode statistics (RecordEpisodeStatistics) -----------------------------
    const bool ended = terminated || truncated;
    if (p.mode == 0 && lead) {
        if (stepping) {
            const double er = dadd(p.st.ep_return[e], reward);
            const int el = p.st.ep_length[e] + 1;
            p.st.ep_return[e] = er;
            p.st.ep_length[e] = el;
            if (p.io.ep_mask) out_store(p, p.io.ep_mask, (size_t)e, (uint8_t)ended);
            if (p.io.ep_return) out_store(p, p.io.ep_return, (size_t)e, ended ? er : 0.0);
            if (p.io.ep_length) out_store(p, p.io.ep_length, (size_t)e, (int32_t)(ended ? el : 0));
            if (p.io.ep_stats && ended) {
                atomicAdd(p.io.ep_stats + 0, er);
                atomicAdd(p.io.ep_stats + 1, (double)el);
                atomicAdd(p.io.ep_stats + 2, 1.0);
            }
        } else {
            if (p.io.ep_mask) out_store(p, p.io.ep_mask, (size_t)e, (uint8_t)0);
            if (p.io.ep_return) out_store(p, p.io.ep_return, (size_t)e, 0.0);
            if (p.io.ep_length) out_store(p, p.io.ep_length, (size_t)e, (int32_t)0);
        }
    }
    // per-car info of the step itself (before any same-step reset)
    if (p.mode == 0 && is_car) {
        if (p.io.info_f64) {
            double* o = p.io.info_f64 + 5 * ci;
            o[0] = x; o[1] = y;
            o[2] = sqrt(dadd(dmul(vx, vx), dmul(vy, vy)));
            o[3] = (flags & F_FINISHED) ? 1.0 : ddiv((double)pidx, Nd);
            o[4] = delta;
        }
        if (p.io.info_i32) {
            int32_t* o = p.io.info_i32 + 4 * ci;
            o[0] = (flags & F_CRASHED) != 0; o[1] = (flags & F_FINISHED) != 0;
            o[2] = placement; o[3] = pidx;
        }
    }

    // ---- reset (racing_env.py:86-102 / multi_racing_env.py:118-153) ------------------
    if (p.mode == 0 && p.autoreset == RK_AUTORESET_SAME_STEP && ended) resetting = true;
    if (resetting) {
        x = tmp->start_x; y = tmp->start_y; ang = tmp->start_angle;  // car.py:17-24
        if (KIND == RK_ENV_MULTI) {  // multi:124-138
            const int slot = p.io.start_slot ? p.io.start_slot[c]
                                             : philox_start_slot(p.seed, e, p.st.reset_count[e], A, a);
            const double center = ddiv((double)(A - 1), 2.0);
            const double off = dmul(dsub((double)slot, center), 3.5);
            x = dadd(tmp->start_x, dmul(tmp->start_nx, off));
            y = dadd(tmp->start_y, dmul(tmp->start_ny, off));
        }
        vx = 0.0; vy = 0.0; last_steer = 0.f;
        pidx = 0; lpidx = 0; flags = 0; fstep = 0; steps = 0;
    }
    __syncwarp();  // every lane has read reset_count before the lead lane bumps it
    if (resetting && lead) {
        p.st.ep_return[e] = 0.0;
        p.st.ep_length[e] = 0;
        p.st.reset_count[e] += 1;
    }

    // ---- write state back ----------------------------------------------------------------
    // a crashed car of a multi env is frozen from now on: this launch leaves its wall distances in wall_cache (below)
    const bool cache_in = KIND == RK_ENV_MULTI && QUERY == RK_QUERY_CULLED && p.mode == 0 && !resetting &&
                          (flags & F_RAYCACHE) && (flags & F_CRASHED);
    if (KIND == RK_ENV_MULTI && QUERY == RK_QUERY_CULLED && p.mode == 0 && p.io.obs != nullptr && is_car && (flags & F_CRASHED))
        flags |= F_RAYCACHE;
    if (p.mode != 2 && is_car) {
        p.st.x[c] = x; p.st.y[c] = y; p.st.ang[c] = ang; p.st.vx[c] = vx; p.st.vy[c] = vy;
        p.st.last_steer[c] = last_steer;
        p.st.pidx[c] = pidx; p.st.lpidx[c] = lpidx; p.st.flags[c] = flags; p.st.fstep[c] = fstep;
        if (a == 0) {
            p.st.steps[e] = steps;
            if (p.mode == 0)
                p.st.needs_reset[e] = (p.autoreset == RK_AUTORESET_NEXT_STEP) ? (stepping && ended) : 0;
            else if (resetting)
                p.st.needs_reset[e] = 0;
        }
    }
    if (p.mode == 0 && is_car) {
        const double r = stepping ? reward : 0.0;
        if (p.io.reward_f32) p.io.reward_f32[ci] = (float)r;
        if (p.io.reward_f64) out_store(p, p.io.reward_f64, (size_t)ci, r);
        if (a == 0) {
            out_store(p, p.io.terminated, (size_t)e, (uint8_t)terminated);
            out_store(p, p.io.truncated, (size_t)e, (uint8_t)truncated);
            if (p.io.done) p.io.done[e] = ended;
            if (p.io.done_f32) p.io.done_f32[e] = ended ? 1.f : 0.f;
        }
    }
    float* obs = p.io.obs;
    if (obs == nullptr) return;
    const bool want_obs = is_car && (p.mode != 1 || resetting);

    // ---- observations (racing_env.py:44-75 / multi_racing_env.py:48-105) ------------------
    { const double2 sc = sincos_d(ang); sn = sc.x; cs = sc.y; }
    __syncwarp();
    S.x[lane] = x; S.y[lane] = y; S.c[lane] = cs; S.s[lane] = sn; S.vx[lane] = vx; S.vy[lane] = vy;
    {
        const double lx[4] = {2.0, 2.0, -2.0, -2.0}, ly[4] = {1.0, -1.0, -1.0, 1.0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            S.cx[k][lane] = dadd(dadd(dmul(cs, lx[k]), dmul(-sn, ly[k])), x);
            S.cy[k][lane] = dadd(dadd(dmul(sn, lx[k]), dmul(cs, ly[k])), y);
        }
    }
    __syncwarp();
    // non-ray slots, every car in parallel
    float nrv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // car 0's non-ray values, for the zero-copy host row
    if (want_obs) {
        float* orow = obs + ci * D + R;
        const double vf = clipd(ddiv(dadd(dmul(vx, cs), dmul(vy, sn)), kMaxSpeed), -1.0, 1.0);
        const double vl = clipd(ddiv(dadd(dmul(-vx, sn), dmul(vy, cs)), kMaxSpeed), -1.0, 1.0);
        orow[0] = (float)vf;
        orow[1] = (float)vl;
        orow[2] = 0.f;  // Car.angular_velocity is never updated
        orow[3] = last_steer;
        nrv[0] = (float)vf; nrv[1] = (float)vl; nrv[2] = 0.f; nrv[3] = last_steer;
        if (KIND == RK_ENV_MULTI) {
            const double mtd = tmp->max_track_distance;
            int w = 4;
            for (int o = 0; o < A; ++o) {
                if (o == a) continue;
                const int l = base + o;
                const double rx = dsub(S.x[l], x), ry = dsub(S.y[l], y);
                const double rvx = dsub(S.vx[l], vx), rvy = dsub(S.vy[l], vy);
                const float o0 = (float)clipd(ddiv(dadd(dmul(rx, cs), dmul(ry, sn)), mtd), -1.0, 1.0);
                const float o1 = (float)clipd(ddiv(dadd(dmul(-rx, sn), dmul(ry, cs)), mtd), -1.0, 1.0);
                const float o2 = (float)clipd(ddiv(dadd(dmul(rvx, cs), dmul(rvy, sn)), kMaxSpeed), -1.0, 1.0);
                const float o3 = (float)clipd(ddiv(dadd(dmul(-rvx, sn), dmul(rvy, cs)), kMaxSpeed), -1.0, 1.0);
                if (w == 4) { nrv[4] = o0; nrv[5] = o1; nrv[6] = o2; nrv[7] = o3; }
                orow[w++] = o0; orow[w++] = o1; orow[w++] = o2; orow[w++] = o3;
            }
        }
    }

    // zero-copy rows: S.vx / S.vy are dead from here on; their 128 floats become the per-environment store of car 0's
    // 4A non-ray values ([4A][32/A] floats: 4 x 32 single-car, 8 x 16 two-car), picked up by the lanes that write the
    // complete host row after the raycast
    float* nr_sh = reinterpret_cast<float*>(S.vx);
    const int nr_stride = 32 / A;   // >= epw
    if (p.obs_host0 != nullptr) {
        __syncwarp();   // every lane has finished reading S.vx / S.vy
        if (want_obs && a == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < 4 * A) nr_sh[k * nr_stride + g] = nrv[k];
        }
        __syncwarp();
    }

    if (STAGED && p.mode != 0) {  // reset / observe launches did not pass the wait above
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // rays: the warp walks over its environments, all lanes cooperating on one
    const unsigned obs_envs = __ballot_sync(kFull, want_obs && a == 0);
    const int nslot = A * R;
    if (QUERY != RK_QUERY_EXACT_F64) {
        unsigned short* win_sh = cv.list + p.list_cap;   // [32][R]: fp32 winner of every ray of the warp's cars (culled mode)
        if (QUERY == RK_QUERY_CULLED) {
            // ---- candidate search, environment by environment, all lanes cooperating (angular sweep) ------------------
            for (int gg = 0; gg < n_env; ++gg) {
                const int gbase = gg * A;
                if (!((obs_envs >> gbase) & 1u)) continue;
                const TrackMeta* tme = STAGED ? &stm : tp.meta + __shfl_sync(kFull, tid, gbase);
                // fp32 ray directions by angle addition from the car's (cos, sin): one lane per (car, ray) slot
                for (int s0 = 0; s0 < nslot; s0 += 32) {
                    const int slot = s0 + lane;
                    if (slot < nslot) {
                        const int ca = slot / R, r = slot - ca * R;
                        const float cc = (float)S.c[gbase + ca], ss = (float)S.s[gbase + ca];
                        const float rc = (float)p.sensor_cos[r], rs = (float)p.sensor_sin[r];
                        cv.dir32[slot] = make_float2(cc * rc - ss * rs, ss * rc + cc * rs);
                        cv.ray_key[slot] = kNoKey;
                    }
                }
                __syncwarp();
                for (int ca = 0; ca < A; ++ca)
                    if (!__shfl_sync(kFull, (int)cache_in, gbase + ca))   // (a frozen car's walls are in wall_cache)
                        raycast_walls_culled<KIND>(tp, tme, p, S.x[gbase + ca], S.y[gbase + ca], S.c[gbase + ca],
                                                   S.s[gbase + ca], ca * R, lane, cv);
                for (int s0 = 0; s0 < nslot; s0 += 32) {
                    const int slot = s0 + lane;
                    if (slot < nslot) {
                        const unsigned long long key = cv.ray_key[slot];
                        win_sh[gbase * R + slot] = key != kNoKey ? (unsigned short)(key & 0xffffu) : (unsigned short)0xffffu;
                    }
                }
                __syncwarp();
            }
        }
        float* row_sh = cv.rows;   // car 0's rays of every environment, for the host rows (the sweep's keys are dead by then)
        if (QUERY == RK_QUERY_CULLED) {
            // ---- float64 distances, one lane per (environment, car, ray) slot of the WHOLE warp: the re-evaluation of the
            //      fp32 winners, the other cars' edges and the stores run at full width whatever the number of cars per
            //      warp (small batches give a warp 2 environments: 44 slots are 2 passes, not 11 at 4 lanes)
            const int total = n_env * nslot;
            const float inv_nslot = 1.f / (float)nslot, inv_R = 1.f / (float)R;
#pragma unroll 1
            for (int s0 = 0; s0 < total; s0 += 32) {
                const int slot = min(s0 + lane, total - 1);
                const int gg = (int)(((float)slot + 0.5f) * inv_nslot);          // exact for these small integers
                const int rem = slot - gg * nslot, ca = (int)(((float)rem + 0.5f) * inv_R), r = rem - ca * R;
                const int lc = gg * A + ca;                                        // the lane that holds this ray's car
                const bool live = s0 + lane < total && ((obs_envs >> (gg * A)) & 1u);
                const int tidr = __shfl_sync(kFull, tid, lc);
                const TrackMeta* tmr = STAGED ? &stm : tp.meta + tidr;
                const double ox = S.x[lc], oy = S.y[lc];
                double v3x = 0.0, v3y = 1.0, wall = INFINITY;
                bool redo = false;
                const bool cached = KIND == RK_ENV_MULTI && __shfl_sync(kFull, (int)cache_in, lc);
                const int flags_c = (KIND == RK_ENV_MULTI) ? __shfl_sync(kFull, flags, lc) : 0;
                float* wcache = (KIND == RK_ENV_MULTI) ? p.st.wall_cache + (size_t)__shfl_sync(kFull, c, lc) * R + r : nullptr;
                if (live) {
                    const double cc = S.c[lc], ss = S.s[lc], rc = p.sensor_cos[r], rs = p.sensor_sin[r];
                    v3x = -dadd(dmul(ss, rc), dmul(cc, rs)); v3y = dsub(dmul(cc, rc), dmul(ss, rs));  // track.py:178
                    const unsigned w = cached ? 0xffffu : win_sh[lc * R + r];
                    if (cached) wall = (double)*wcache;
                    if (w != 0xffffu) {
                        const size_t i = 2 * (size_t)tmr->wp_off + w;
                        const double ax = tp.v2x[i], ay = tp.v2y[i];
                        const double v1x = dsub(ox, tp.sx[i]), v1y = dsub(oy, tp.sy[i]);
                        wall = ray_segment<true>(v1x, v1y, ax, ay, dsub(dmul(ax, v1y), dmul(ay, v1x)), v3x, v3y, kWallMinDot);
                        redo = wall == INFINITY;  // fp32 candidate rejected by the float64 test
                    }
                }
                unsigned fb = __ballot_sync(kFull, redo);
                while (fb) {  // rare: that ray against every wall, exactly, with the whole warp
                    const int b = __ffs(fb) - 1;
                    fb &= fb - 1;
                    const TrackMeta tmb = STAGED ? stm : tp.meta[__shfl_sync(kFull, tidr, b)];
                    const double t = raycast_wall_exact_one(tp, tmb, __shfl_sync(kFull, ox, b), __shfl_sync(kFull, oy, b),
                                                            __shfl_sync(kFull, v3x, b), __shfl_sync(kFull, v3y, b), lane);
                    if (lane == b) wall = t;
                }
                if (live) {
                    double t = wall;
                    // (the float copy loses nothing: rounding is monotonic, so min and the final float cast commute)
                    if (KIND == RK_ENV_MULTI && (flags_c & F_RAYCACHE) && !cached) *wcache = (float)wall;
                    if (KIND == RK_ENV_MULTI)
                        t = fmin(fmin(t, raycast_car_edges<true>(S, gg * A, A, ox, oy, v3x, v3y)), kMaxRange);  // multi_track.py:8,26
                    else if (t == INFINITY)
                        t = kMaxRange;  // track.py:196-197
                    const int ee = STAGED ? genv[gg] : e_base + gg;
                    const size_t oi = agent_major ? (size_t)ca * p.E + ee : (size_t)ee * A + ca;
                    const float hval = __fdiv_rn((float)t, 50.0f);  // racing_env.py:46-53
                    obs[oi * D + r] = hval;
                    if (p.obs_host0 != nullptr && ca == 0) row_sh[gg * R + r] = hval;
                }
            }
        }
        // ---- grid mode: every lane casts and finishes the R rays of ITS car, one after the other --------------------
        const TrackMeta* tmr = tmp;
        const float o32x = (float)(x - tmr->org_x), o32y = (float)(y - tmr->org_y);
        const double* sx = tp.sx + 2 * (size_t)tmr->wp_off;
        const double* sy = tp.sy + 2 * (size_t)tmr->wp_off;
        const double* v2x = tp.v2x + 2 * (size_t)tmr->wp_off;
        const double* v2y = tp.v2y + 2 * (size_t)tmr->wp_off;
        float* orow = obs + ci * D;
        // Grid mode: pass j works on each car's j-th LONGEST ray of the previous step (p.st.ray_order): the lanes of a
        // warp then walk rays of similar length at the same time and the traversal loops stay converged.  Only a
        // schedule: every ray is cast exactly once whatever the order.
        unsigned long long order = (QUERY == RK_QUERY_GRID && is_car && R <= 15) ? p.st.ray_order[c] : 0ull;
        const bool ordered = (order >> 60) == 0xFull;
#pragma unroll 1
        for (int j = 0; QUERY == RK_QUERY_GRID && j < R; ++j) {
            const int r = ordered ? min((int)((order >> (4 * j)) & 15ull), R - 1) : j;
            double v3x = 0.0, v3y = 1.0, wall = INFINITY;
            bool redo = false;
            if (want_obs) {
                // ray direction by angle addition from the car's (cos, sin)
                const double rc = p.sensor_cos[r], rs = p.sensor_sin[r];
                const double dcs = dsub(dmul(cs, rc), dmul(sn, rs)), dsn = dadd(dmul(sn, rc), dmul(cs, rs));
                v3x = -dsn; v3y = dcs;  // track.py:178
                const RayHit hit = grid_ray(tp, tmr, o32x, o32y, (float)dcs, (float)dsn, (KIND == RK_ENV_MULTI) ? 50.01f : INFINITY);
                redo = !hit.inside;
                if (hit.s0 >= 0) {
                    {
                        const int i = hit.s0;
                        const double ax = v2x[i], ay = v2y[i];
                        const double v1x = dsub(x, sx[i]), v1y = dsub(y, sy[i]);
                        wall = ray_segment<true>(v1x, v1y, ax, ay, dsub(dmul(ax, v1y), dmul(ay, v1x)), v3x, v3y, kWallMinDot);
                    }
                    // the runner-up decides when float64 rejects the fp32 winner or the two are within fp32 rounding
                    const bool close = hit.s1 >= 0 && hit.t1 <= hit.t0 + 2e-4f * (1.f + hit.t0);
                    if (hit.s1 >= 0 && (wall == INFINITY || close)) {
                        const int i = hit.s1;
                        const double ax = v2x[i], ay = v2y[i];
                        const double v1x = dsub(x, sx[i]), v1y = dsub(y, sy[i]);
                        wall = fmin(wall, ray_segment<true>(v1x, v1y, ax, ay, dsub(dmul(ax, v1y), dmul(ay, v1x)), v3x, v3y, kWallMinDot));
                    }
                    redo = redo || wall == INFINITY;  // fp32 candidates rejected by the float64 test
                }
            }
            unsigned fb = __ballot_sync(kFull, redo);
            while (fb) {  // rare: that ray against every wall, exactly, with the whole warp
                const int b = __ffs(fb) - 1;
                fb &= fb - 1;
                const TrackMeta tmb = tp.meta[__shfl_sync(kFull, tid, b)];
                const double t = raycast_wall_exact_one(tp, tmb, __shfl_sync(kFull, x, b), __shfl_sync(kFull, y, b),
                                                        __shfl_sync(kFull, v3x, b), __shfl_sync(kFull, v3y, b), lane);
                if (lane == b) wall = t;
            }
            if (want_obs) {
                double t = wall;
                if (KIND == RK_ENV_MULTI)
                    t = fmin(fmin(t, raycast_car_edges<true>(S, base, A, x, y, v3x, v3y)), kMaxRange);  // multi_track.py:8,26
                else if (t == INFINITY)
                    t = kMaxRange;  // track.py:196-197
                const float hval = __fdiv_rn((float)t, 50.0f);  // racing_env.py:46-53
                orow[r] = hval;
                if (p.obs_host0 != nullptr && a == 0) row_sh[g * R + r] = hval;
            }
        }
        if (QUERY == RK_QUERY_GRID && want_obs && R <= 15 && p.mode != 2) {
            // next step's order: rank the fresh readings (re-read from the row just written: static indexing)
            float hv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) hv[k] = k < R ? orow[k] : -1.f;
            unsigned long long next = 0xFull << 60;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                int rank = 0;
#pragma unroll
                for (int m = 0; m < 16; ++m) rank += (hv[m] > hv[k] || (hv[m] == hv[k] && m < k)) ? 1 : 0;
                if (k < R) next |= (unsigned long long)k << (4 * rank);
            }
            p.st.ray_order[c] = next;
        }
        if (p.obs_host0 != nullptr) {
            // car 0's complete rows of the warp's consecutive environments form ONE contiguous run of n_env * D floats
            // in the caller's pinned buffer: full-width coalesced stores
            __syncwarp();
            const int run = n_env * D;
            for (int f0 = 0; f0 < run; f0 += 32) {
                const int f = f0 + lane;
                if (f < run) {
                    const int gg = f / D, col = f - gg * D;
                    if ((obs_envs >> (gg * A)) & 1u)
                        p.obs_host0[(size_t)(e_base + gg) * D + col] = col < R ? row_sh[gg * R + col] : nr_sh[(col - R) * nr_stride + gg];
                }
            }
        }
        return;
    }
    for (int gg = 0; gg < n_env; ++gg) {
        const int gbase = gg * A;
        if (!((obs_envs >> gbase) & 1u)) continue;
        const int ee = STAGED ? genv[gg] : e_base + gg;
        const TrackMeta tm = STAGED ? stm : tp.meta[__shfl_sync(kFull, tid, gbase)];
        {
            for (int ca = 0; ca < A; ++ca) {
                const double ox = S.x[gbase + ca], oy = S.y[gbase + ca];
                const double oang = __shfl_sync(kFull, ang, gbase + ca);
                const size_t oi = agent_major ? (size_t)ca * p.E + ee : (size_t)ee * A + ca;
                float* orow = obs + oi * D;
                for (int r0 = 0; r0 < R; r0 += kRayBlock) {
                    const int nr = min(kRayBlock, R - r0);
                    double v3x[kRayBlock], v3y[kRayBlock], best[kRayBlock];
#pragma unroll
                    for (int k = 0; k < kRayBlock; ++k) {
                        // lane k computes ray r0+k's direction, then it is broadcast
                        double dsn = 0.0, dcs = 1.0;
                        if (lane == k && k < nr) { const double2 sc = sincos_d(dadd(oang, p.sensor_angles[r0 + k])); dsn = sc.x; dcs = sc.y; }
                        v3x[k] = -__shfl_sync(kFull, dsn, k);  // track.py:178 v3 = (-dir_y, dir_x)
                        v3y[k] = __shfl_sync(kFull, dcs, k);
                        best[k] = INFINITY;
                    }
                    raycast_walls_exact(tp, tm, ox, oy, v3x, v3y, nr, lane, best);
#pragma unroll
                    for (int k = 0; k < kRayBlock; ++k) {
                        double t = warp_min_d(best[k]);
                        if (lane == k && k < nr) {
                            if (KIND == RK_ENV_MULTI)
                                t = fmin(fmin(t, raycast_car_edges(S, gbase, A, ox, oy, v3x[k], v3y[k])), kMaxRange);
                            else if (t == INFINITY)
                                t = kMaxRange;
                            orow[r0 + k] = __fdiv_rn((float)t, 50.0f);
                        }
                    }
                }
            }
        }
    }
}

}  // namespace

int launch_step(const StepParams& p, int query_mode, int env_kind, cudaStream_t stream) {
    const bool staged = p.group_env != nullptr;
    const int epw = p.epw;
    const int warps = (p.env_end - p.env_begin + epw - 1) / epw;
    const int grid = staged ? p.n_ctas : (warps + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid <= 0) return 0;
    const size_t smem = kWarpsPerCta * warp_smem_bytes(p.A, p.R, p.list_cap) + (staged ? (size_t)p.stage_bytes : 0);
    using Kern = void (*)(const StepParams);
    const bool single = env_kind == RK_ENV_SINGLE, culled = query_mode == RK_QUERY_CULLED;
    Kern k;
    if (query_mode == RK_QUERY_GRID)
        k = single ? step_kernel<RK_ENV_SINGLE, RK_QUERY_GRID, false> : step_kernel<RK_ENV_MULTI, RK_QUERY_GRID, false>;
    else if (staged && culled)
        k = single ? step_kernel<RK_ENV_SINGLE, RK_QUERY_CULLED, true> : step_kernel<RK_ENV_MULTI, RK_QUERY_CULLED, true>;
    else if (culled)
        k = single ? step_kernel<RK_ENV_SINGLE, RK_QUERY_CULLED, false> : step_kernel<RK_ENV_MULTI, RK_QUERY_CULLED, false>;
    else
        k = single ? step_kernel<RK_ENV_SINGLE, RK_QUERY_EXACT_F64, false> : step_kernel<RK_ENV_MULTI, RK_QUERY_EXACT_F64, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    StepParams q = p;
    if (!(staged && culled)) q.group_env = nullptr;
    k<<<grid, kWarpsPerCta * 32, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace rk
