// K7: projection of units onto the border polyline (replaces the per-unit LibGEOS calls of projection_points,
// src/border_projection.jl:4-22: distance(border, point) and nearestPoints(border, point)[1], which feed
// proj_estimator :24-31 and weights_proj src/weight_at_units.jl:44-53).
//
// GEOS 3.8 semantics restated (algorithm/Distance.cpp pointToSegment, geom/LineSegment.cpp closestPoint, DistanceOp's
// scan over the segments of every line in order, keeping the FIRST minimum):
//   A == B            : d = |p - A|
//   r = ((p-A).(B-A)) / |B-A|^2 ;  r <= 0 : d = |p - A| ;  r >= 1 : d = |p - B|
//   otherwise         : s = ((A.y-p.y)(B.x-A.x) - (A.x-p.x)(B.y-A.y)) / |B-A|^2 ;  d = |s| sqrt(|B-A|^2)
//   closest point     : 0 < r < 1 ? A + r (B - A) : (|p-A| < |p-B| ? A : B)
// Every operation is a separately rounded IEEE operation (__dmul_rn / __dadd_rn / ..., no FMA contraction), so the
// result -- including which units fall within maxdist -- is bit-identical to the CPU restatement.
//
// HBM-bound by design and tiny: one thread per unit, the polyline is staged through shared memory in tiles of 1024
// vertices (coalesced 16-byte loads), n x nv segment tests.
#include "kernels.h"

namespace geordd {

constexpr int GEOM_TILE = 1024;

__device__ __forceinline__ double dist_rn(double px, double py, double ax, double ay) {
    double dx = __dsub_rn(px, ax), dy = __dsub_rn(py, ay);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

__global__ void __launch_bounds__(256) projection_kernel(const double2* __restrict__ X, int n, const double2* __restrict__ V,
                                                         const unsigned char* __restrict__ brk, int nv,
                                                         double2* __restrict__ Xproj, double* __restrict__ dist) {
    __shared__ double2 sv[GEOM_TILE + 1];
    __shared__ unsigned char sb[GEOM_TILE + 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const double2 p = live ? X[i] : make_double2(0.0, 0.0);
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    double2 bA = make_double2(0.0, 0.0), bB = bA;
    for (int v0 = 0; v0 < nv - 1; v0 += GEOM_TILE) {
        const int cnt = min(GEOM_TILE + 1, nv - v0);   // vertices v0 .. v0+cnt-1 -> cnt-1 segments
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            sv[t] = V[v0 + t];
            sb[t] = brk[v0 + t];
        }
        __syncthreads();
        if (live) {
            for (int s = 0; s + 1 < cnt; ++s) {
                if (sb[s + 1]) continue;   // vertex s+1 starts a new line of the MultiLineString: no segment
                const double2 A = sv[s], B = sv[s + 1];
                double d;
                if (A.x == B.x && A.y == B.y) {
                    d = dist_rn(p.x, p.y, A.x, A.y);
                } else {
                    const double ex = __dsub_rn(B.x, A.x), ey = __dsub_rn(B.y, A.y);
                    const double len2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                    const double r = __ddiv_rn(__dadd_rn(__dmul_rn(__dsub_rn(p.x, A.x), ex), __dmul_rn(__dsub_rn(p.y, A.y), ey)), len2);
                    if (r <= 0.0) d = dist_rn(p.x, p.y, A.x, A.y);
                    else if (r >= 1.0) d = dist_rn(p.x, p.y, B.x, B.y);
                    else {
                        const double sn = __dsub_rn(__dmul_rn(__dsub_rn(A.y, p.y), ex), __dmul_rn(__dsub_rn(A.x, p.x), ey));
                        d = __dmul_rn(fabs(__ddiv_rn(sn, len2)), __dsqrt_rn(len2));
                    }
                }
                if (d < best) { best = d; bA = A; bB = B; }
            }
        }
    }
    if (!live) return;
    // closest point on the winning segment (LineSegment::closestPoint)
    double2 q;
    double r;
    if (p.x == bA.x && p.y == bA.y) r = 0.0;
    else if (p.x == bB.x && p.y == bB.y) r = 1.0;
    else {
        const double ex = __dsub_rn(bB.x, bA.x), ey = __dsub_rn(bB.y, bA.y);
        const double len2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
        r = __ddiv_rn(__dadd_rn(__dmul_rn(__dsub_rn(p.x, bA.x), ex), __dmul_rn(__dsub_rn(p.y, bA.y), ey)), len2);
    }
    if (r > 0.0 && r < 1.0) {
        q.x = __dadd_rn(bA.x, __dmul_rn(r, __dsub_rn(bB.x, bA.x)));
        q.y = __dadd_rn(bA.y, __dmul_rn(r, __dsub_rn(bB.y, bA.y)));
    } else {
        q = dist_rn(p.x, p.y, bA.x, bA.y) < dist_rn(p.x, p.y, bB.x, bB.y) ? bA : bB;
    }
    Xproj[i] = q;
    dist[i] = best;
}

// Point-in-region for the grid of infinite_proj_sentinels (LibGEOS.within(p, region), src/border_projection.jl:70):
// GEOS RayCrossingCounter restated per ring segment (p1, p2) for the test point q --
//   both ends strictly left of q: skip;  q == p2: on boundary;  horizontal segment at q.y: on boundary if q.x within it;
//   ((p1.y > q.y && p2.y <= q.y) || (p2.y > q.y && p1.y <= q.y)): orientation sign of (p1, p2, q), flipped if p2.y < p1.y,
//   0 -> on boundary, left turn -> one crossing
// -- with the orientation determinant in plain (separately rounded) FP64 instead of GEOS's double-double; the two differ
// only for points within rounding of an edge.  Crossing parities are XOR-ed over ALL rings (shells and holes of every
// polygon: even-odd rule == interior of a valid MultiPolygon); boundary points are not within.
__global__ void __launch_bounds__(256) in_region_kernel(const double2* __restrict__ P, int n, const double2* __restrict__ R,
                                                        const unsigned char* __restrict__ brk, int nr, unsigned char* __restrict__ inside) {
    __shared__ double2 sv[GEOM_TILE + 1];
    __shared__ unsigned char sb[GEOM_TILE + 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const double2 q = live ? P[i] : make_double2(0.0, 0.0);
    int crossings = 0;
    bool on_boundary = false;
    for (int v0 = 0; v0 < nr - 1; v0 += GEOM_TILE) {
        const int cnt = min(GEOM_TILE + 1, nr - v0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            sv[t] = R[v0 + t];
            sb[t] = brk[v0 + t];
        }
        __syncthreads();
        if (live) {
            for (int s = 0; s + 1 < cnt; ++s) {
                if (sb[s + 1]) continue;   // vertex s+1 starts a new ring
                const double2 p1 = sv[s], p2 = sv[s + 1];
                if (p1.x < q.x && p2.x < q.x) continue;
                if (q.x == p2.x && q.y == p2.y) { on_boundary = true; continue; }
                if (p1.y == q.y && p2.y == q.y) {
                    const double minx = fmin(p1.x, p2.x), maxx = fmax(p1.x, p2.x);
                    if (q.x >= minx && q.x <= maxx) on_boundary = true;
                    continue;
                }
                if ((p1.y > q.y && p2.y <= q.y) || (p2.y > q.y && p1.y <= q.y)) {
                    const double det = __dsub_rn(__dmul_rn(__dsub_rn(p2.x, p1.x), __dsub_rn(q.y, p1.y)),
                                                 __dmul_rn(__dsub_rn(p2.y, p1.y), __dsub_rn(q.x, p1.x)));
                    int sign = det > 0.0 ? 1 : (det < 0.0 ? -1 : 0);
                    if (sign == 0) { on_boundary = true; continue; }
                    if (p2.y < p1.y) sign = -sign;
                    if (sign == 1) ++crossings;
                }
            }
        }
    }
    if (live) inside[i] = (!on_boundary && (crossings & 1)) ? 1 : 0;
}

void launch_in_region(cudaStream_t st, const double* P, int n, const double* R, const unsigned char* brk, int nr,
                      unsigned char* inside) {
    in_region_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double2*>(P), n, reinterpret_cast<const double2*>(R), brk, nr,
                                                    inside); ++g_kernel_launches;
}

void launch_projection(cudaStream_t st, const double* X, int n, const double* V, const unsigned char* brk, int nv,
                       double* Xproj, double* dist) {
    projection_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double2*>(X), n, reinterpret_cast<const double2*>(V), brk,
                                                     nv, reinterpret_cast<double2*>(Xproj), dist); ++g_kernel_launches;
}

}  // namespace geordd
