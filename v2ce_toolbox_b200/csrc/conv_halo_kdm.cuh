// Halo-tile Conv3d (3x3x3, stride 1, pad 1) with the three depth taps merged into the MMA's N dimension, for
// layers with few output channels (N tile = 32).
//
// Why: tcgen05.mma with both operands in shared memory reads 128 rows of A and N rows of B (32 bytes each) per
// K=16 step; the shared-memory port moves 128 B/clk, so one step costs max(N/2, (128+N)/4) cycles
// (tools/mma_rate.py on B200: N=32 44.7, N=64 50.2, N=128 64, N=256 128).  At N=32 the tensor pipe is therefore
// busy 36 % of the time at best: the A operand is re-read for every 32 output channels.
//
// Input slice z under depth tap kd contributes to output slice z-kd+1 AT THE SAME accumulator rows.  With the
// accumulators of consecutive output slices in consecutive TMEM column blocks, ONE instruction of N = 3*32 = 96
//     D[:, (z-1 .. z+1) * 32 ..] += A_z(tap kh,kw) x [W(kd=2); W(kd=1); W(kd=0)]^T
// does the work of three, for one read of A: 57 cycles instead of 134.  Slices outside the CTA's depth block
// are dropped by shrinking N and offsetting the B descriptor by whole 32-row blocks; the first touch of an
// accumulator block (output slice z+1 at input slice z) is one extra N=32 instruction without the accumulate flag.
//
// Loop order per tile (spatial tile x T=8 output slices x 32 channels): channel chunk -> input slice -> 9 taps.
// The nine [3 kd][32][64] weight tiles of a chunk stay resident in shared memory (108 KB) while the T+2 input
// patches stream through a TMA ring; every patch is multiplied exactly once.  Depth slices -1 and D are zero
// padding and are skipped.  Two accumulator sets (2 x 8 x 32 = 512 TMEM columns): the epilogue of tile i
// overlaps the MMAs of tile i+1, and runs on 8 warps (two per TMEM lane quarter, alternate slices) because at
// this MMA rate the epilogue would otherwise set the pace.
//
// Layers with 64 output channels run as two N tiles of 32 (112 cycles per three taps instead of 150).
// Replaces the same reference calls as conv_halo.cuh (submodules.py:249-263).
#pragma once
#include "conv_halo.cuh"

namespace v2ce {
namespace halo {

constexpr int kKdmThreads = 384;      // warp 0 TMA patches, warp 1 weight tiles, warp 2 MMA issue, warp 3 idle, warps 4-11 epilogue
constexpr int kKdmBN = 32;
constexpr int kKdmT = 8;
constexpr int kKdmWTile = 3 * kKdmBN * kBlockK * 2;   // bytes of one (chunk, tap) weight tile: [3][32][64] bf16

__global__ void __launch_bounds__(kKdmThreads) conv_halo_kdm_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                     const __grid_constant__ CUtensorMap tm1,
                                                                     const HaloArgs a) {
  constexpr int BN = kKdmBN, T = kKdmT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  const int n_tiles = a.Cout / BN;
  const int total_tiles = a.B * (a.D / T) * a.tiles_h * a.tiles_w * n_tiles;
  const int ncc = a.ncc0 + a.ncc1;

  const uint32_t w_base = base;                                   // 9 resident weight tiles
  const uint32_t a_base = base + 9u * kKdmWTile;                  // patch ring
  const uint32_t bar_base = a_base + (uint32_t)a.SA * a.a_stage_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxSA + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * kMaxSA + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * kMaxSA + 9 + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (2 * kMaxSA + 18 + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (2 * kMaxSA + 20 + s); };
  const int bar_bytes = (2 * kMaxSA + 22) * 8;
  uint8_t* tail = smem + (bar_base - base) + bar_bytes;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(tail);
  float* s_scale = reinterpret_cast<float*>(tail + 16);
  float* s_shift = s_scale + BN;
  float* s_pred = s_shift + BN;                   // [20][32] weights + [20] bias (used when a.pred_w != nullptr)

  if (tid == 0) {
    for (int s = 0; s < a.SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < 9; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr))),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_ptr;

  // input slices a tile multiplies: d0-1 .. d0+T, minus the zero-padding slices -1 and D
  auto z_lo = [&](int d0) { return d0 > 0 ? d0 - 1 : 0; };
  auto z_hi = [&](int d0) { return d0 + T < a.D ? d0 + T : a.D - 1; };

  if (warp == 0) {
    // ================= patch producer (TMA) =================
    int s = 0, ph = 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, a, T, n_tiles);
      const int zl = z_lo(tc.d0), zh = z_hi(tc.d0);
      for (int cc = 0; cc < ncc; ++cc) {
        const bool first = cc < a.ncc0;
        const CUtensorMap* map = first ? &tm0 : &tm1;
        const int c0 = (first ? cc : cc - a.ncc0) * kBlockK;
        for (int z = zl; z <= zh; ++z) {
          mbar_wait(a_empty(s), (uint32_t)ph, a.error_flag);
          if (elect_one()) {
            mbar_arrive_expect_tx(a_full(s), (uint32_t)a.box_bytes);
            tma_load_5d(a_base + (uint32_t)s * a.a_stage_bytes, map, c0, tc.w0 - 1, tc.h0 - 1, z, tc.b, a_full(s));
          }
          __syncwarp();
          if (++s == a.SA) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= weight-tile producer: slot = tap, refilled once per (tile, chunk) =================
    uint32_t ph = 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, a, T, n_tiles);
      for (int cc = 0; cc < ncc; ++cc) {
        const __nv_bfloat16* wt = a.wpack + (size_t)(tc.n_tile * ncc + cc) * 9 * (kKdmWTile / 2);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(w_empty(tap), ph, a.error_flag);
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full(tap), (uint32_t)kKdmWTile);
            bulk_copy_g2s(w_base + (uint32_t)tap * kKdmWTile, wt + (size_t)tap * (kKdmWTile / 2), (uint32_t)kKdmWTile, w_full(tap));
          }
          __syncwarp();
        }
        ph ^= 1u;
      }
    }
  } else if (warp == 2) {
    // ================= MMA issuer =================
    int sa = 0, pa = 0;                             // patch ring position / parity
    uint32_t pw = 0;                                // parity of the weight slots for the current (tile, chunk)
    const uint32_t pw8 = (uint32_t)a.PW * 8u;       // one patch row of pixels in descriptor units (16 B)
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const TileCoord tc = decode_tile(tile, a, T, n_tiles);
      const int zl = z_lo(tc.d0), zh = z_hi(tc.d0);
      const int ab = iter & 1, use = iter >> 1;
      mbar_wait(tmem_empty(ab), (uint32_t)((use & 1) ^ 1), a.error_flag);     // epilogue has drained this accumulator set
      tcgen05_fence_after();
      const uint32_t acc_base = tmem_acc + (uint32_t)(ab * T * BN);
      for (int cc = 0; cc < ncc; ++cc) {
        // zero-padded channels (64-channel pitch of a 32-channel tensor) are not multiplied
        const int rem = (cc < a.ncc0) ? (a.real0 - cc * kBlockK) : (a.real1 - (cc - a.ncc0) * kBlockK);
        const int ks = rem >= kBlockK ? kBlockK / 16 : (rem + 15) / 16;
#pragma unroll 1
        for (int z = zl; z <= zh; ++z) {
          mbar_wait(a_full(sa), (uint32_t)pa, a.error_flag);
          tcgen05_fence_after();
          const uint32_t patch = a_base + (uint32_t)sa * a.a_stage_bytes;
          // output slices z-1+j, j = 0..2 (depth tap kd = 2-j), clipped to the block [d0, d0+T)
          const int jlo = (z - 1 >= tc.d0) ? 0 : (tc.d0 - z + 1);
          const int jhi = (z + 1 <= tc.d0 + T - 1) ? 2 : (tc.d0 + T - z);
          const uint32_t col = acc_base + (uint32_t)((z - 1 + jlo - tc.d0) * BN);
          const uint32_t n_all = (uint32_t)((jhi - jlo + 1) * BN);
          const uint32_t idesc_all = make_idesc((int)n_all);
          // descriptor low words of (tap 0, k 0); a tap adds (kh*PW + kw) rows of 128 B to A and one weight tile to
          // B, a K=16 step adds 32 B to both (fully unrolled: the issue loop must stay well under the 56 cycles
          // one N=96 instruction occupies the tensor pipe)
          const uint32_t a_lo = smem_desc_lo(patch);
          const uint32_t b_lo = smem_desc_lo(w_base + (uint32_t)jlo * (BN * 128));
          const bool first_z = (z == zl), last_z = (z == zh);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (first_z) {
              mbar_wait(w_full(tap), pw, a.error_flag);
              tcgen05_fence_after();
            }
            const uint32_t a_t = a_lo + (uint32_t)(tap / 3) * pw8 + (uint32_t)(tap % 3) * 8u;
            const uint32_t b_t = b_lo + (uint32_t)tap * (kKdmWTile >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (k < ks) {
                if (tap != 0 || k != 0 || cc != 0) {
                  tcgen05_mma_bf16_lo(col, a_t + 2 * k, b_t + 2 * k, idesc_all, 1u);
                } else if (first_z) {
                  // first multiply into this tile's accumulators: nothing covered has been written yet
                  tcgen05_mma_bf16_lo(col, a_t, b_t, idesc_all, 0u);
                } else {
                  // slice z+1 (j = 2) is touched for the first time, the others accumulate
                  const int j_old_hi = jhi < 1 ? jhi : 1;
                  const uint32_t n_old = (uint32_t)((j_old_hi - jlo + 1) * BN);
                  tcgen05_mma_bf16_lo(col, a_t, b_t, make_idesc((int)n_old), 1u);
                  if (jhi == 2) tcgen05_mma_bf16_lo(col + n_old, a_t, b_t + n_old * 8u, make_idesc(BN), 0u);
                }
              }
            }
            if (last_z) tcgen05_commit_elect(w_empty(tap));     // last use of this chunk's tap tile
          }
          tcgen05_commit_elect(a_empty(sa));
          if (++sa == a.SA) { sa = 0; pa ^= 1; }
        }
        pw ^= 1u;
      }
      tcgen05_commit_elect(tmem_full(ab));
    }
  } else if (warp >= 4) {
    // ================= epilogue (warps 4-11) =================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;             // even / odd slices of the block
    const int i = quarter * 32 + lane;            // accumulator row == TMEM lane
    const int th = i / a.PW, tw = i % a.PW;
    const int etid = tid - 128;                   // 0..255 inside the epilogue group
    const float isg = a.inv_sigma ? __ldg(a.inv_sigma) : 1.f;
    if (a.pred_w != nullptr) {
      for (int j = etid; j < 660; j += 256) s_pred[j] = j < 640 ? __ldg(a.pred_w + j) : __ldg(a.pred_b + (j - 640));
    }
    int cur_n_tile = -1;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const TileCoord tc = decode_tile(tile, a, T, n_tiles);
      if (tc.n_tile != cur_n_tile) {              // (re)stage the folded BatchNorm scale/shift of this channel tile
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int j = etid; j < BN; j += 256) {
          s_scale[j] = __ldg(a.scale + tc.n_tile * BN + j) * isg;
          s_shift[j] = __ldg(a.shift + tc.n_tile * BN + j);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_n_tile = tc.n_tile;
      }
      const int ab = iter & 1, use = iter >> 1;
      mbar_wait(tmem_full(ab), (uint32_t)(use & 1), a.error_flag);
      __syncwarp();
      tcgen05_fence_after();
      halo_epilogue_slices<BN>(a, tc, n_tiles, tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * T * BN), half, 2, T,
                               th, tw, s_scale, s_shift, s_pred);
      // this accumulator set may be overwritten by the MMAs of a later tile
      tcgen05_fence_before();
      mbar_arrive(tmem_empty(ab));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512u) : "memory");
  }
}

// fp32 (Cout, Cin_real, 3,3,3) -> bf16 [Cout/32][cc][kh*3+kw][j = 2-kd][32][64] (rows swizzled); channel padding as in
// pack_weights_halo_kernel
__global__ void pack_weights_kdm_kernel(const float* __restrict__ w, int Cout, int cin_real, int pad0, int real0, int pad1,
                                        int real1, __nv_bfloat16* __restrict__ out) {
  constexpr int BN = kKdmBN;
  const int ncc = (pad0 + pad1) / kBlockK;
  const size_t total = (size_t)Cout * 27 * ncc * kBlockK;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % 8);
    const int qs = (int)((i / 8) % 8);
    const int r = (int)((i / 64) % BN);
    size_t tile = i / ((size_t)64 * BN);        // ((n_tile*ncc + cc)*9 + tap9)*3 + j
    const int j = (int)(tile % 3); tile /= 3;
    const int tap9 = (int)(tile % 9); tile /= 9;
    const int cc = (int)(tile % ncc);
    const int n_tile = (int)(tile / ncc);
    const int kd = 2 - j;
    const int q = qs ^ (r & 7);
    const int p = cc * kBlockK + q * 8 + e;      // padded input channel
    int c = -1;
    if (p < pad0) { if (p < real0) c = p; }
    else { const int p1 = p - pad0; if (p1 < real1) c = real0 + p1; }
    const int n = n_tile * BN + r;
    float v = 0.f;
    if (c >= 0) v = w[((size_t)n * cin_real + c) * 27 + kd * 9 + tap9];
    out[i] = __float2bfloat16_rn(v);
  }
}

struct KdmPlan {
  TileShape ts;
  int SA, a_stage_bytes, box_bytes, smem_bytes;
  bool ok;
};

// the kernel applies to depth multiples of 8 and needs room for at least 3 patch stages beside the weights
inline KdmPlan plan_kdm(int D, int H, int W) {
  KdmPlan p;
  p.ts = pick_tile(H, W);
  const int rows = p.ts.PW * (p.ts.TH + 2);
  p.box_bytes = rows * 128;
  p.a_stage_bytes = ((rows + 2 + 7) / 8) * 1024;
  const int tail = (2 * kMaxSA + 22) * 8 + 16 + 2 * kKdmBN * 4 + 660 * 4;
  const int budget = 227 * 1024 - 1024 - tail - 9 * kKdmWTile;
  p.SA = budget / p.a_stage_bytes;
  if (p.SA > kMaxSA) p.SA = kMaxSA;
  p.ok = (D % kKdmT == 0) && p.SA >= 3;
  p.smem_bytes = 9 * kKdmWTile + p.SA * p.a_stage_bytes + tail + 1024;
  return p;
}

inline int launch_halo_kdm(const CUtensorMap& tm0, const CUtensorMap& tm1, const HaloArgs& a, int smem_bytes, cudaStream_t s) {
  static int configured = 0;
  if (configured < smem_bytes) {
    V2CE_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kdm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = smem_bytes;
  }
  if (a.T != kKdmT || a.D % kKdmT != 0 || a.Cout % kKdmBN != 0)
    return set_error(V2CE_ERR_INVALID, "depth-merged halo kernel: depth %d / Cout %d not supported", a.D, a.Cout);
  const int total = a.B * (a.D / kKdmT) * a.tiles_h * a.tiles_w * (a.Cout / kKdmBN);
  const int grid = total < sm_count_cached() ? total : sm_count_cached();
  conv_halo_kdm_kernel<<<grid, kKdmThreads, smem_bytes, s>>>(tm0, tm1, a);
  V2CE_LAUNCH_CHECK("conv_halo_kdm_kernel");
  return V2CE_OK;
}

}  // namespace halo
}  // namespace v2ce
