// Image pre-processing on the device for frames that are not at the model's resolution.
// Replaces /root/reference/v2ce.py:45-64 (image_pre_processing): /255 in float32, cv2.resize (INTER_LINEAR) to the
// model's height, pair stacking, Normalize(0.153, 0.165).  Frames already at the model's height never come here:
// for them the resize is the identity and the rest is fused into the head conv (unet.cu, head_prep_kernel<true>).
//
// Arithmetic contract (oracle/resize_oracle.py, pinned against the installed OpenCV):
//   coordinates   f = (d + 0.5) * (src / dst) - 0.5 in double; s = floor(f); a = float32(f - s);
//                 s < 0 -> (0, a = 0); s >= src - 1 -> (src - 1, a = 0); second tap min(s + 1, src - 1)
//   horizontal    h = fma(S[x1] - S[x0], ax, S[x0])          float32, one rounding for the fma
//   vertical      v = fma(h[y1] - h[y0], ay, h[y0])
//   unit          (v - 0.153f) / 0.165f                      float32 subtract and IEEE divide
// HBM-bound gather: one thread per output pixel of one frame reads four source bytes and writes the value to the
// two image units the frame belongs to (second channel of pair i-1, first channel of pair i).
#include "common.cuh"

namespace v2ce {
namespace preproc {

struct Tap {
  int i0, i1;
  float a;
};

// one rounding per operation, nothing contracted: the double ops go through __d*_rn
__device__ __forceinline__ Tap make_tap(int d, double scale, int src) {
  const double f = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  const double fl = floor(f);
  long long s = (long long)fl;
  float a = __double2float_rn(__dsub_rn(f, fl));
  if (s < 0) { s = 0; a = 0.f; }
  if (s >= src - 1) { s = src - 1; a = 0.f; }
  Tap t;
  t.i0 = (int)s;
  t.i1 = min((int)s + 1, src - 1);
  t.a = a;
  return t;
}

__global__ void __launch_bounds__(256) image_units_kernel(const uint8_t* __restrict__ frames, int frames_per_window,
                                                          int src_h, int src_w, int dst_h, int dst_w, double scale_y,
                                                          double scale_x, float* __restrict__ units) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= dst_h * dst_w) return;
  const int fi = blockIdx.y, win = blockIdx.z;          // frame inside the window, window
  const int dy = pix / dst_w, dx = pix - dy * dst_w;
  const Tap ty = make_tap(dy, scale_y, src_h), tx = make_tap(dx, scale_x, src_w);
  const uint8_t* f = frames + ((size_t)win * frames_per_window + fi) * ((size_t)src_h * src_w);
  const uint8_t* r0 = f + (size_t)ty.i0 * src_w;
  const uint8_t* r1 = f + (size_t)ty.i1 * src_w;
  const float s00 = __fdiv_rn((float)__ldg(r0 + tx.i0), 255.f), s01 = __fdiv_rn((float)__ldg(r0 + tx.i1), 255.f);
  const float s10 = __fdiv_rn((float)__ldg(r1 + tx.i0), 255.f), s11 = __fdiv_rn((float)__ldg(r1 + tx.i1), 255.f);
  const float h0 = __fmaf_rn(__fsub_rn(s01, s00), tx.a, s00);
  const float h1 = __fmaf_rn(__fsub_rn(s11, s10), tx.a, s10);
  const float v = __fmaf_rn(__fsub_rn(h1, h0), ty.a, h0);
  const float u = __fdiv_rn(__fsub_rn(v, 0.153f), 0.165f);
  // units (windows, L, 2, dst_h, dst_w), L = frames_per_window - 1: frame fi is channel 0 of pair fi and channel 1 of pair fi-1
  const int L = frames_per_window - 1;
  const size_t plane = (size_t)dst_h * dst_w;
  float* w0 = units + (size_t)win * L * 2 * plane;
  if (fi < L) w0[((size_t)fi * 2 + 0) * plane + pix] = u;
  if (fi >= 1) w0[((size_t)(fi - 1) * 2 + 1) * plane + pix] = u;
}

}  // namespace preproc
}  // namespace v2ce

using namespace v2ce;

extern "C" int v2ce_image_units(const uint8_t* frames_dev, int32_t n_windows, int32_t frames_per_window, int32_t src_h,
                                int32_t src_w, int32_t dst_h, int32_t dst_w, float* units_dev, void* stream) {
  V2CE_REQUIRE(frames_dev && units_dev, "NULL device pointer");
  V2CE_REQUIRE(n_windows > 0 && n_windows <= 65535 && frames_per_window >= 2 && frames_per_window <= 65535,
               "bad window geometry: %d windows of %d frames", n_windows, frames_per_window);
  V2CE_REQUIRE(src_h >= 2 && src_w >= 1 && dst_h >= 1 && dst_w >= 1, "bad frame geometry %dx%d -> %dx%d", src_h, src_w,
               dst_h, dst_w);      // one source row takes a different path in OpenCV (oracle/resize_oracle.py)
  V2CE_REQUIRE((long long)dst_h * dst_w < (1LL << 31) && (long long)src_h * src_w < (1LL << 31), "frame too large");
  const double scale_y = (double)src_h / (double)dst_h, scale_x = (double)src_w / (double)dst_w;
  dim3 grid((unsigned)(((long long)dst_h * dst_w + 255) / 256), (unsigned)frames_per_window, (unsigned)n_windows);
  preproc::image_units_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames_dev, frames_per_window, src_h, src_w, dst_h, dst_w, scale_y, scale_x, units_dev);
  V2CE_LAUNCH_CHECK("preproc::image_units_kernel");
  return V2CE_OK;
}
