// Event-frame preview on sm_100a: per-pair accumulation of the 10 voxel bins, exact
// percentile by radix select over float bit patterns, float64 normalisation to BGR uint8.
// Replaces /root/reference/v2ce.py:241-280 (write_event_frame_video) up to the cv2 encoder.
//
// All three kernels are HBM-bound streaming passes (8,906,040 B per 346x260 pair).
#include "common.cuh"

namespace v2ce {
namespace ef {

constexpr int kThreads = 256;

// E1 (v2ce.py:255 / 259): s = v0; s += v1; ... sequential fp32 adds, the order numpy uses
// when it reduces a strided axis.  Gray mode reduces (polarity, bin) polarity-major.
template <int V>
__global__ void __launch_bounds__(kThreads) accumulate_kernel(const float* __restrict__ vox, int HW, int keep_polarity,
                                                               float* __restrict__ sums) {
  const int plane = blockIdx.y;                    // keep: n*2+p ; gray: n
  const int pix = (blockIdx.x * kThreads + threadIdx.x) * V;
  if (pix >= HW) return;
  const int nterms = keep_polarity ? 10 : 20;
  const float* src = vox + (size_t)plane * (keep_polarity ? 10 : 20) * HW + pix;
  if (V == 4) {
    float4 s = __ldg(reinterpret_cast<const float4*>(src));
    for (int c = 1; c < nterms; ++c) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * HW));
      s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y); s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
    }
    *reinterpret_cast<float4*>(sums + (size_t)plane * HW + pix) = s;
  } else {
    float s = __ldg(src);
    for (int c = 1; c < nterms; ++c) s = __fadd_rn(s, __ldg(src + (size_t)c * HW));
    sums[(size_t)plane * HW + pix] = s;
  }
}

// E2 (v2ce.py:262-264): radix select.  Positive floats are order-isomorphic to their bit
// patterns, so the k-th smallest positive sum is found with four 8-bit histogram passes.
// One kernel per pass: every block histograms its share into per-warp shared-memory copies, adds them to the global
// histogram of the pass, and the LAST block to finish (atomic ticket) picks the bucket of both ranks and extends the
// prefixes for the next pass -- round 1 ran a single-thread kernel between the passes (7-34 us each, pure latency;
// profiles/ncu_ef_r2_a.txt) and read the sums with scalar loads (1.8 TB/s).
struct SelectState {
  unsigned long long hist[4][2][256];
  unsigned long long npos;
  unsigned long long remaining[2];   // rank still to skip inside the current prefix bucket
  unsigned int prefix[2];
  unsigned int done[4];              // blocks that have added their histogram of pass p
};

// block-wide (256 threads): first bucket d in [0, 255) whose inclusive count exceeds k (255 if none), and the count
// before it -- what a serial walk over the 256 buckets yields.
__device__ __forceinline__ void pick_bucket(const unsigned long long* h, unsigned long long k,
                                            unsigned long long* sh_warp /*8*/, unsigned long long* sh_out /*2*/) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned long long c = __ldcg(h + t);          // written by other blocks' atomics in this launch: L2, not L1
  unsigned long long inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sh_warp[warp] = inc;
  if (t == 0) { sh_out[0] = 255ull; sh_out[1] = ~0ull; }
  __syncthreads();
  unsigned long long off = 0ull;
  for (int w = 0; w < warp; ++w) off += sh_warp[w];
  inc += off;
  const unsigned long long exc = inc - c;
  if (t < 255 && inc > k && exc <= k) { sh_out[0] = (unsigned long long)t; sh_out[1] = exc; }     // exactly one thread
  if (t == 255) sh_warp[8] = exc;                       // count before bucket 255 (the fall-through case)
  __syncthreads();
}
// result of the last pick_bucket (any thread, after its barrier): the count before the picked bucket.  The fall-through
// value is kept in its own word instead of being patched into sh_out by thread 0: compute-sanitizer's racecheck
// flagged that patch (the compiler hoists the predicate's load to every thread while thread 0 stores).
__device__ __forceinline__ unsigned long long picked_before(const unsigned long long* sh_warp, const unsigned long long* sh_out) {
  return sh_out[1] == ~0ull ? sh_warp[8] : sh_out[1];
}

__global__ void __launch_bounds__(kThreads) select_hist_kernel(const float* __restrict__ v, long long n, int pass,
                                                                double percentile, int multiplicity, SelectState* st,
                                                                long long* __restrict__ result) {
  __shared__ unsigned int h[kThreads / 32][2][256];     // per-warp copies: 8x less contention on a hot bucket
  __shared__ unsigned long long sh_warp[9], sh_out[2];
  __shared__ unsigned int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kThreads / 32) * 512; i += kThreads) (&h[0][0][0])[i] = 0u;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned int p0 = st->prefix[0], p1 = st->prefix[1];

  auto add = [&](float x, bool in_range) {
    const unsigned int b = __float_as_uint(x);
    const bool pos = in_range && x > 0.f;
    const unsigned int d = (b >> shift) & 255u;
    const unsigned int hi = (pass == 0) ? 0u : (b >> (shift + 8));
    const bool m0 = pos && (pass == 0 || hi == p0), m1 = pos && pass != 0 && hi == p1;
    // the top byte of small positive floats is one or two exponent buckets: aggregate a warp that agrees
    const unsigned vm0 = __ballot_sync(0xffffffffu, m0);
    if (vm0) {
      const int first = __ffs(vm0) - 1;
      const unsigned d0 = __shfl_sync(0xffffffffu, d, first);
      if (__all_sync(0xffffffffu, !m0 || d == d0)) { if (lane == first) atomicAdd(&h[warp][0][d0], (unsigned)__popc(vm0)); }
      else if (m0) atomicAdd(&h[warp][0][d], 1u);
    }
    const unsigned vm1 = __ballot_sync(0xffffffffu, m1);
    if (vm1) {
      const int first = __ffs(vm1) - 1;
      const unsigned d0 = __shfl_sync(0xffffffffu, d, first);
      if (__all_sync(0xffffffffu, !m1 || d == d0)) { if (lane == first) atomicAdd(&h[warp][1][d0], (unsigned)__popc(vm1)); }
      else if (m1) atomicAdd(&h[warp][1][d], 1u);
    }
  };

  const long long stride = (long long)gridDim.x * kThreads;
  const bool vec = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  // whole warps iterate together (the ballots above need every lane): round the trip count up per warp
  const long long first4 = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long trips4 = (n4 + stride - 1) / stride;
  for (long long it = 0; it < trips4; ++it) {
    const long long i = first4 + it * stride;
    const bool ok = i < n4;
    const float4 x = ok ? __ldg(reinterpret_cast<const float4*>(v) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    add(x.x, ok); add(x.y, ok); add(x.z, ok); add(x.w, ok);
  }
  const long long tail0 = n4 * 4;
  const long long tripst = (n - tail0 + stride - 1) / stride;
  for (long long it = 0; it < tripst; ++it) {
    const long long i = tail0 + first4 + it * stride;
    const bool ok = i < n;
    add(ok ? __ldg(v + i) : 0.f, ok);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += kThreads) {
    unsigned int c = 0u;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) c += (&h[w][0][0])[i];
    if (c) atomicAdd(&st->hist[pass][0][0] + i, (unsigned long long)c);
  }
  // ---- the last block to arrive closes the pass ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->done[pass], 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned long long* h0 = &st->hist[pass][0][0];
  const unsigned long long* h1 = (pass == 0) ? h0 : &st->hist[pass][1][0];       // pass 0: both ranks share one histogram
  if (pass == 0) {
    // total count of positive values, then numpy's 'linear' method: virtual index (n-1)*q with q = percentile/100 (float64)
    pick_bucket(h0, ~0ull - 1ull, sh_warp, sh_out);                              // never found: sh_out[1] = count before bucket 255
    if (threadIdx.x == 0) {
      const unsigned long long npos = picked_before(sh_warp, sh_out) + __ldcg(h0 + 255);
      st->npos = npos;
      result[0] = (long long)npos;
      if (npos == 0) {
        result[1] = 0; result[2] = 0; result[3] = 0;
      } else {
        const double q = __ddiv_rn(percentile, 100.0);
        const double nn = (double)(npos * (unsigned long long)multiplicity);
        const double vi = __dmul_rn(nn - 1.0, q);
        long long lo = (long long)floor(vi);
        const long long last = (long long)(npos * (unsigned long long)multiplicity) - 1;
        if (lo < 0) lo = 0;
        if (lo > last) lo = last;
        const long long hi = lo + 1 > last ? last : lo + 1;
        st->remaining[0] = (unsigned long long)(lo / multiplicity);
        st->remaining[1] = (unsigned long long)(hi / multiplicity);
        result[1] = lo;
      }
    }
    __syncthreads();
  }
  if (st->npos == 0) return;
  for (int r = 0; r < 2; ++r) {
    pick_bucket(r == 0 ? h0 : h1, st->remaining[r], sh_warp, sh_out);
    if (threadIdx.x == 0) {
      st->remaining[r] -= picked_before(sh_warp, sh_out);
      st->prefix[r] = (st->prefix[r] << 8) | (unsigned int)sh_out[0];
    }
    __syncthreads();
  }
  if (pass == 3 && threadIdx.x == 0) {
    result[2] = (long long)st->prefix[0];
    result[3] = (long long)st->prefix[1];
  }
}

// E3 (v2ce.py:267-277): clip(s,0,ub)/ub in float64, *255, truncate to uint8; channels
// (R,G,B) = (pos, neg, 0) then RGB->BGR, i.e. stored order B=0, G=neg, R=pos.
__device__ __forceinline__ unsigned int norm_u8(float s, double ub) {
  double x = (double)s;
  x = x < 0.0 ? 0.0 : x;
  x = x > ub ? ub : x;
  x = __dmul_rn(__ddiv_rn(x, ub), 255.0);
  return (unsigned int)(unsigned char)x;
}

template <int V>
__global__ void __launch_bounds__(kThreads) normalize_kernel(const float* __restrict__ sums, int HW, int keep_polarity,
                                                              double ub, uint8_t* __restrict__ frames) {
  const int n = blockIdx.y;
  const int pix = (blockIdx.x * kThreads + threadIdx.x) * V;
  if (pix >= HW) return;
  unsigned int bytes[V * 3];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if (keep_polarity) {
      const float pos = __ldg(sums + ((size_t)n * 2 + 0) * HW + pix + v);
      const float neg = __ldg(sums + ((size_t)n * 2 + 1) * HW + pix + v);
      bytes[v * 3 + 0] = 0u;
      bytes[v * 3 + 1] = norm_u8(neg, ub);
      bytes[v * 3 + 2] = norm_u8(pos, ub);
    } else {
      const unsigned int g = norm_u8(__ldg(sums + (size_t)n * HW + pix + v), ub);
      bytes[v * 3 + 0] = g; bytes[v * 3 + 1] = g; bytes[v * 3 + 2] = g;
    }
  }
  uint8_t* dst = frames + ((size_t)n * HW + pix) * 3;
  if (V == 4) {
    uint32_t w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      w[i] = bytes[i * 4] | (bytes[i * 4 + 1] << 8) | (bytes[i * 4 + 2] << 16) | (bytes[i * 4 + 3] << 24);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    d32[0] = w[0]; d32[1] = w[1]; d32[2] = w[2];
  } else {
#pragma unroll
    for (int i = 0; i < V * 3; ++i) dst[i] = (uint8_t)bytes[i];
  }
}

}  // namespace ef
}  // namespace v2ce

using namespace v2ce;
using namespace v2ce::ef;

extern "C" int v2ce_ef_accumulate(const float* voxels_dev, int32_t n_pairs, int32_t height, int32_t width,
                                  int32_t keep_polarity, float* sums_dev, void* stream) {
  V2CE_REQUIRE(voxels_dev && sums_dev, "NULL device pointer");
  V2CE_REQUIRE(n_pairs > 0 && height > 0 && width > 0, "bad geometry");
  const int HW = height * width;
  const int planes = keep_polarity ? n_pairs * 2 : n_pairs;
  V2CE_REQUIRE(planes <= 65535, "too many pairs per call (max %d)", keep_polarity ? 32767 : 65535);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (HW % 4 == 0) {
    dim3 grid((HW / 4 + kThreads - 1) / kThreads, planes);
    accumulate_kernel<4><<<grid, kThreads, 0, s>>>(voxels_dev, HW, keep_polarity, sums_dev);
  } else {
    dim3 grid((HW + kThreads - 1) / kThreads, planes);
    accumulate_kernel<1><<<grid, kThreads, 0, s>>>(voxels_dev, HW, keep_polarity, sums_dev);
  }
  V2CE_LAUNCH_CHECK("ef::accumulate_kernel");
  return V2CE_OK;
}

extern "C" int v2ce_ef_select_workspace_bytes(size_t* bytes) {
  V2CE_REQUIRE(bytes != nullptr, "bytes is NULL");
  *bytes = align_up(sizeof(SelectState), 256);
  return V2CE_OK;
}

extern "C" int v2ce_ef_select(const float* sums_dev, int64_t n_values, double percentile, int32_t multiplicity,
                              void* ws_dev, size_t ws_bytes, int64_t* result_dev, void* stream) {
  V2CE_REQUIRE(sums_dev && ws_dev && result_dev, "NULL device pointer");
  V2CE_REQUIRE(n_values > 0 && multiplicity >= 1, "bad arguments");
  V2CE_REQUIRE(percentile >= 0.0 && percentile <= 100.0, "percentile must be in [0,100]");
  if (ws_bytes < sizeof(SelectState))
    return set_error(V2CE_ERR_WORKSPACE, "select workspace too small: need %zu, got %zu", sizeof(SelectState), ws_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SelectState* st = static_cast<SelectState*>(ws_dev);
  V2CE_CUDA_CHECK(cudaMemsetAsync(st, 0, sizeof(SelectState), s));
  long long want = (n_values / 4 + kThreads * 4 - 1) / (kThreads * 4);
  const int cap = sm_count_cached() * 8;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  for (int pass = 0; pass < 4; ++pass) {
    select_hist_kernel<<<grid, kThreads, 0, s>>>(sums_dev, n_values, pass, percentile, multiplicity, st,
                                                 reinterpret_cast<long long*>(result_dev));
    V2CE_LAUNCH_CHECK("ef::select_hist_kernel");
  }
  return V2CE_OK;
}

extern "C" int v2ce_ef_normalize(const float* sums_dev, int32_t n_pairs, int32_t height, int32_t width,
                                 int32_t keep_polarity, double upper_bound, uint8_t* frames_dev, void* stream) {
  V2CE_REQUIRE(sums_dev && frames_dev, "NULL device pointer");
  V2CE_REQUIRE(n_pairs > 0 && n_pairs <= 65535 && height > 0 && width > 0, "bad geometry");
  V2CE_REQUIRE(upper_bound > 0.0, "upper_bound must be positive");
  const int HW = height * width;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (HW % 4 == 0) {
    dim3 grid((HW / 4 + kThreads - 1) / kThreads, n_pairs);
    normalize_kernel<4><<<grid, kThreads, 0, s>>>(sums_dev, HW, keep_polarity, upper_bound, frames_dev);
  } else {
    dim3 grid((HW + kThreads - 1) / kThreads, n_pairs);
    normalize_kernel<1><<<grid, kThreads, 0, s>>>(sums_dev, HW, keep_polarity, upper_bound, frames_dev);
  }
  V2CE_LAUNCH_CHECK("ef::normalize_kernel");
  return V2CE_OK;
}
