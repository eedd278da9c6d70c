// History colours along weighted-ensemble lineages + the coloured transition records for K3.
//
// reference: NonMarkovModel.fit (msm_we/nmm.py:117-167) walks every discrete trajectory: a frame's colour is A / B when
// its state is in A / B, otherwise the colour of the previous frame (undefined until the trajectory first touches A or
// B; the first frame is never coloured: the walk starts at index `lag`), and every transition whose two colours are
// defined adds 1 to C[2 s_prev + col_prev, 2 s_now + col_now].  For WE data the trajectories are the lineages of the
// walkers of the last iteration, traced back through seg_index['parent_id']; lineages share their early segments, so a
// segment's transition is counted once per surviving descendant.  Here that is three passes over the segments instead
// of a Python loop over (leaf x depth):
//   colour : forward over iterations, colour_now[s] = rule(label_now[s], colour_prev[parent[s]])
//   leaves : backward over iterations, leaves_prev[parent[s]] += leaves_now[s]      (integer atomics: exact)
//   records: per segment (s_prev, s_now, col_prev, col_now, weight = leaves_now[s]) -> K3 with C = 2.
#include "common.cuh"

namespace mwe {

__global__ void __launch_bounds__(256)
    lineage_colour_kernel(const int64_t* __restrict__ label_now, const int64_t* __restrict__ parent, int64_t S_now,
                          const int8_t* __restrict__ colour_prev, int64_t S_prev, const uint8_t* __restrict__ state_class,
                          int64_t n_states, int8_t* __restrict__ colour_now) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_now) return;
    const int64_t s = label_now[i];
    const uint8_t cls = (s >= 0 && s < n_states) ? state_class[s] : 0;      // 1 = in A, 2 = in B
    int8_t c = -1;
    if (cls == 1) c = 0;
    else if (cls == 2) c = 1;
    else if (colour_prev) {
        const int64_t p = parent[i];
        if (p >= 0 && p < S_prev) c = colour_prev[p];
    }
    colour_now[i] = c;
}

__global__ void __launch_bounds__(256)
    lineage_leaves_kernel(const int64_t* __restrict__ parent, const unsigned long long* __restrict__ leaves_now, int64_t S_now,
                          unsigned long long* __restrict__ leaves_prev, int64_t S_prev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_now) return;
    const int64_t p = parent[i];
    const unsigned long long n = leaves_now[i];
    if (n && p >= 0 && p < S_prev) atomicAdd(leaves_prev + p, n);
}

__global__ void __launch_bounds__(256)
    lineage_records_kernel(const int64_t* __restrict__ label_prev, const int8_t* __restrict__ colour_prev, int64_t S_prev,
                           const int64_t* __restrict__ label_now, const int8_t* __restrict__ colour_now,
                           const int64_t* __restrict__ parent, const unsigned long long* __restrict__ leaves_now, int64_t S_now,
                           int64_t* __restrict__ start, int64_t* __restrict__ end, uint8_t* __restrict__ col0,
                           uint8_t* __restrict__ col1, double* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S_now) return;
    const int64_t p = parent[i];
    int64_t s0 = 0, s1 = 0;
    uint8_t c0 = 0, c1 = 0;
    double weight = 0.0;                       // uncoloured / parentless transitions contribute nothing
    if (p >= 0 && p < S_prev) {
        const int8_t a = colour_prev[p], b = colour_now[i];
        if (a >= 0 && b >= 0) {
            s0 = label_prev[p]; s1 = label_now[i];
            c0 = (uint8_t)a; c1 = (uint8_t)b;
            weight = (double)leaves_now[i];
        }
    }
    start[i] = s0; end[i] = s1; col0[i] = c0; col1[i] = c1; w[i] = weight;
}

}  // namespace mwe

extern "C" int mwe_lineage_colour(const int64_t* label_now, const int64_t* parent, int64_t S_now, const int8_t* colour_prev,
                                  int64_t S_prev, const uint8_t* state_class, int64_t n_states, int8_t* colour_now, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(S_now >= 0 && S_prev >= 0 && n_states >= 0, "lineage_colour: bad sizes");
    if (S_now == 0) return MWE_OK;
    MWE_REQUIRE(label_now && state_class && colour_now && (parent || !colour_prev), "lineage_colour: null pointer");
    lineage_colour_kernel<<<(unsigned)((S_now + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        label_now, parent, S_now, colour_prev, S_prev, state_class, n_states, colour_now);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_lineage_leaves(const int64_t* parent, const uint64_t* leaves_now, int64_t S_now, uint64_t* leaves_prev,
                                  int64_t S_prev, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(S_now >= 0 && S_prev >= 0, "lineage_leaves: bad sizes");
    if (S_now == 0) return MWE_OK;
    MWE_REQUIRE(parent && leaves_now && leaves_prev, "lineage_leaves: null pointer");
    lineage_leaves_kernel<<<(unsigned)((S_now + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        parent, reinterpret_cast<const unsigned long long*>(leaves_now), S_now, reinterpret_cast<unsigned long long*>(leaves_prev), S_prev);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}

extern "C" int mwe_lineage_records(const int64_t* label_prev, const int8_t* colour_prev, int64_t S_prev, const int64_t* label_now,
                                   const int8_t* colour_now, const int64_t* parent, const uint64_t* leaves_now, int64_t S_now,
                                   int64_t* start, int64_t* end, uint8_t* col0, uint8_t* col1, double* w, void* stream) {
    using namespace mwe;
    MWE_REQUIRE(S_now >= 0 && S_prev >= 0, "lineage_records: bad sizes");
    if (S_now == 0) return MWE_OK;
    MWE_REQUIRE(label_prev && colour_prev && label_now && colour_now && parent && leaves_now && start && end && col0 && col1 && w,
                "lineage_records: null pointer");
    lineage_records_kernel<<<(unsigned)((S_now + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        label_prev, colour_prev, S_prev, label_now, colour_now, parent, reinterpret_cast<const unsigned long long*>(leaves_now), S_now,
        start, end, col0, col1, w);
    MWE_CHECK_LAUNCH();
    return MWE_OK;
}
