// rowgeom.h -- the fixed "virtual chunk" geometry of every row reduction on the path.
//
// north_star asks for the same outer/inner iteration counts as the reference; on the GPU that additionally requires the
// trajectory not to depend on how many GPUs share the rows.  Every quantity that is a sum over residual rows
// (J'(Jv), ||Jv||^2, J'r, ||r||^2 and the two scalars of the incremental Cauchy search) is therefore computed on a geometry
// that depends ONLY on (M_total, n):
//
//   * the M_total rows are cut into kGroups = 8 groups of G chunks each (chunk c = g*G + b covers a contiguous, balanced
//     range of global rows); G = min(148, max(1, M_total / 512));
//   * a chunk is reduced by ONE CTA with a fixed thread / team pattern that only depends on chunk-local row indices
//     => one partial per (chunk, team);
//   * partials are combined by a fixed tree: teams in order, then the G chunks of a group (kChains interleaved chains,
//     combined in order), then the 8 group sums in order.
//
// With N in {1, 2, 4, 8} ranks, rank r owns the groups [r*8/N, (r+1)*8/N): at N = 1 a CTA walks its 8 chunks one after the
// other, at N = 8 it owns one.  Only group sums cross NVLink, and every rank adds the same 8 vectors in the same order, so
// all results are bit-identical at N = 1, 2, 4, 8 (and across ranks).
#pragma once
#include <cuda_runtime.h>

namespace bnl {

constexpr int kGroups = 8;    // virtual ranks: the unit that is exchanged between GPUs
constexpr int kMaxG = 148;    // chunks per group (= CTAs of the persistent row kernels; one per B200 SM)
constexpr int kChains = 8;    // interleaved summation chains over the chunks of a group (fixed tree)

struct RowGeom {
    long long M_total;  // global number of residual rows
    long long row0;     // global index of this rank's first row (= chunk_begin(g0 * G))
    int G;              // chunks per group
    int g0, ng;         // this rank's groups [g0, g0 + ng)

    __host__ __device__ long long chunk_begin(long long c) const {  // global row index, c in [0, kGroups*G]
        const long long nchunk = (long long)kGroups * G;
        const long long base = M_total / nchunk, extra = M_total % nchunk;
        return c * base + (c < extra ? c : extra);
    }
    // local row range of the chunk (gi, b), gi in [0, ng), b in [0, G)
    __host__ __device__ long long local_begin(int gi, int b) const { return chunk_begin((long long)(g0 + gi) * G + b) - row0; }
    __host__ __device__ long long local_end(int gi, int b) const { return chunk_begin((long long)(g0 + gi) * G + b + 1) - row0; }
    __host__ __device__ long long local_rows() const { return chunk_begin((long long)(g0 + ng) * G) - row0; }
};

inline int geom_pick_G(long long M_total) {
    long long g = M_total / 512;
    if (g < 1) g = 1;
    if (g > kMaxG) g = kMaxG;
    return (int)g;
}

// nranks must divide kGroups.  Returns false otherwise.
inline bool geom_make(long long M_total, int nranks, int rank, RowGeom* out) {
    if (nranks < 1 || kGroups % nranks != 0 || rank < 0 || rank >= nranks || M_total < 0) return false;
    RowGeom g{};
    g.M_total = M_total;
    g.G = geom_pick_G(M_total);
    g.ng = kGroups / nranks;
    g.g0 = rank * g.ng;
    g.row0 = 0;
    g.row0 = g.chunk_begin((long long)g.g0 * g.G);
    *out = g;
    return true;
}

}  // namespace bnl
