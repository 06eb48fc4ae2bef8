// cv::resize(src, dst, Size(w, h)) with the default INTER_LINEAR, the way System::TrackRGBD applies it to the colour image
// and to the depth map before Tracking::GrabImageRGBD sees them (reference src/System.cc:325-329), batched over frames.
// Arithmetic = OpenCV's own (modules/imgproc/src/resize.cpp, resizeGeneric_ with HResizeLinear / VResizeLinear):
//   8U (1, 3 or 4 channels): 11-bit fixed-point coefficients, T = s0*c0 + s1*c1, dst = ((b0*(T0>>4))>>16 + (b1*(T1>>4))>>16 + 2)>>2
//     — the model the pyramid uses, per channel;
//   16U / 32F: float coefficients 1-f and f, t = s0*a0 + s1*a1 and d = t0*b0 + t1*b1 with every product and sum rounded
//     on its own (no FMA), 16U stored through saturate_cast<ushort>(cvRound(d)); an exact 2 x 2 reduction is INTER_AREA in
//     cv::resize (the rounded integer mean for 16U; the 8U and 32F bilinear expressions give the area values anyway).
// Bit-identical to cv2 4.13 for 8U (IPP on or off) and, with IPP off, for 16U / 32F (tests/test_resize_oracle.py); OpenCV builds
// that route 16U / 32F through IPP's own resize kernel differ from OpenCV's arithmetic in the last bit — declared in DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "drfe_internal.h"

namespace drfe {

struct ResizeTab { int i0, i1; float a0, a1; int c0, c1; };   // source indices, float and 11-bit coefficients of one destination index

template <typename T, int CH>
__global__ void __launch_bounds__(256) k_resize_linear(const T* __restrict__ src, long long src_rs, long long src_fs, T* __restrict__ dst, long long dst_rs,
                                                       long long dst_fs, int dw, int dh, const ResizeTab* __restrict__ xt, const ResizeTab* __restrict__ yt, bool area2) {
  const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, f = blockIdx.z;
  if (x >= dw || y >= dh) return;
  const ResizeTab tx = xt[x], ty = yt[y];
  const T* r0 = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(src) + f * src_fs + ty.i0 * src_rs);
  const T* r1 = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(src) + f * src_fs + ty.i1 * src_rs);
  T* o = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(dst) + f * dst_fs + y * dst_rs) + (long long)x * CH;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const T s00 = r0[tx.i0 * CH + c], s01 = r0[tx.i1 * CH + c], s10 = r1[tx.i0 * CH + c], s11 = r1[tx.i1 * CH + c];
    if (sizeof(T) == 1) {
      const int T0 = (int)s00 * tx.c0 + (int)s01 * tx.c1, T1 = (int)s10 * tx.c0 + (int)s11 * tx.c1;
      o[c] = (T)((((ty.c0 * (T0 >> 4)) >> 16) + ((ty.c1 * (T1 >> 4)) >> 16) + 2) >> 2);
    } else {
      const float t0 = __fadd_rn(__fmul_rn((float)s00, tx.a0), __fmul_rn((float)s01, tx.a1));
      const float t1 = __fadd_rn(__fmul_rn((float)s10, tx.a0), __fmul_rn((float)s11, tx.a1));
      const float d = __fadd_rn(__fmul_rn(t0, ty.a0), __fmul_rn(t1, ty.a1));
      if (sizeof(T) == 2) o[c] = area2 ? (T)(((int)s00 + (int)s01 + (int)s10 + (int)s11 + 2) >> 2)   // exact 2 x 2 reduction: cv::resize switches to INTER_AREA
                                       : (T)min(max(__float2int_rn(d), 0), 65535);                    // saturate_cast<ushort>(float): cvRound, then clamp
      else o[c] = (T)d;
    }
  }
}

}  // namespace drfe

using namespace drfe;

struct drfe_resizer {
  int device = 0, sw = 0, sh = 0, dw = 0, dh = 0, max_batch = 0;
  cudaStream_t stream = nullptr;
  ResizeTab* d_xt = nullptr;
  ResizeTab* d_yt = nullptr;
  uint8_t* d_src = nullptr;   // staging of host inputs, max_batch * sw * sh * 4 bytes
  uint8_t* d_dst = nullptr;   // staging of host outputs, max_batch * dw * dh * 4 bytes
};

// OpenCV resize.cpp: scale = 1 / (dsize / ssize); fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx.  Horizontally
// an index outside the image is clamped and its fraction set to 0 (xofs / alpha); vertically only the two row indices are
// clipped (clip(sy0 - ksize2 + 1 + k, 0, ssize.height)) and beta keeps the fraction — it matters when upscaling.
static void resize_table(int ssize, int dsize, bool vertical, std::vector<ResizeTab>& out) {
  const double scale = 1.0 / ((double)dsize / ssize);
  out.resize(dsize);
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    ResizeTab t;
    if (vertical) {
      t.i0 = std::min(std::max(s, 0), ssize - 1); t.i1 = std::min(std::max(s + 1, 0), ssize - 1);
    } else {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
      t.i0 = s; t.i1 = std::min(s + 1, ssize - 1);
    }
    t.a0 = 1.f - f; t.a1 = f;
    t.c0 = (int)lrintf(t.a0 * 2048.f); t.c1 = (int)lrintf(t.a1 * 2048.f);
    out[d] = t;
  }
}

extern "C" {

int drfe_resizer_create(int src_width, int src_height, int dst_width, int dst_height, int max_batch, int device, drfe_resizer** out) {
  if (!out) { set_error("drfe_resizer_create: null argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  if (src_width < 2 || src_height < 2 || dst_width < 1 || dst_height < 1 || max_batch < 1 || src_width > 16384 || src_height > 16384 || dst_width > 16384 ||
      dst_height > 16384) {
    set_error("drfe_resizer_create: invalid sizes"); return DRFE_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("drfe_resizer_create: no CUDA device available (there is no CPU fallback)"); return DRFE_ERR_CUDA; }
  if (device < 0 || device >= ndev) { set_error("drfe_resizer_create: bad device %d", device); return DRFE_ERR_ARG; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error("cudaSetDevice(%d) failed", device); return DRFE_ERR_CUDA; }
  drfe_resizer* h = new drfe_resizer();
  h->device = device; h->sw = src_width; h->sh = src_height; h->dw = dst_width; h->dh = dst_height; h->max_batch = max_batch;
  std::vector<ResizeTab> xt, yt;
  resize_table(src_width, dst_width, false, xt);
  resize_table(src_height, dst_height, true, yt);
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_xt, xt.size() * sizeof(ResizeTab));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_yt, yt.size() * sizeof(ResizeTab));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_xt, xt.data(), xt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_yt, yt.data(), yt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { set_error("drfe_resizer_create: %s", cudaGetErrorString(e)); drfe_resizer_destroy(h); return DRFE_ERR_CUDA; }
  *out = h;
  return DRFE_OK;
}

int drfe_resizer_destroy(drfe_resizer* h) {
  if (!h) return DRFE_OK;
  DeviceScope ds(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_xt); cudaFree(h->d_yt); cudaFree(h->d_src); cudaFree(h->d_dst);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DRFE_OK;
}

void* drfe_resizer_stream(drfe_resizer* h) { return h ? (void*)h->stream : nullptr; }

int drfe_resizer_sync(drfe_resizer* h) {
  if (!h) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}

int drfe_resize(drfe_resizer* h, int nframes, const void* src, int pixel_type, int channels, size_t src_row_stride, size_t src_frame_stride, int src_mem_kind,
                void* dst, size_t dst_row_stride, size_t dst_frame_stride, int dst_mem_kind) {
  NvtxRange nvtx_("drfe_resize");
  if (!h || !src || !dst) { set_error("drfe_resize: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_resize: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const size_t esz = pixel_type == DRFE_PIX_U8 ? 1 : (pixel_type == DRFE_PIX_U16 ? 2 : (pixel_type == DRFE_PIX_F32 ? 4 : 0));
  if (esz == 0 || (esz == 1 && channels != 1 && channels != 3 && channels != 4) || (esz != 1 && channels != 1)) {
    set_error("drfe_resize: pixel type %d with %d channels is not supported (8U: 1, 3, 4 channels; 16U and 32F: 1)", pixel_type, channels); return DRFE_ERR_ARG;
  }
  const size_t sline = (size_t)h->sw * channels * esz, dline = (size_t)h->dw * channels * esz;
  if (src_row_stride < sline || dst_row_stride < dline || (nframes > 1 && (src_frame_stride < src_row_stride * h->sh || dst_frame_stride < dst_row_stride * h->dh)) ||
      (src_row_stride % esz) || (dst_row_stride % esz) || (src_frame_stride % esz) || (dst_frame_stride % esz)) {
    set_error("drfe_resize: bad strides"); return DRFE_ERR_ARG;
  }
  if ((src_mem_kind != DRFE_MEM_HOST && src_mem_kind != DRFE_MEM_DEVICE) || (dst_mem_kind != DRFE_MEM_HOST && dst_mem_kind != DRFE_MEM_DEVICE)) { set_error("drfe_resize: bad mem_kind"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const uint8_t* s = (const uint8_t*)src;
  long long srs = (long long)src_row_stride, sfs = (long long)src_frame_stride;
  if (src_mem_kind == DRFE_MEM_HOST) {
    if (!h->d_src) DRFE_CUDA(cudaMalloc((void**)&h->d_src, (size_t)h->max_batch * h->sw * h->sh * 4));
    for (int f = 0; f < nframes; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync(h->d_src + (size_t)f * sline * h->sh, sline, s + (size_t)f * src_frame_stride, src_row_stride, sline, h->sh, cudaMemcpyHostToDevice, st));
    s = h->d_src; srs = (long long)sline; sfs = (long long)(sline * h->sh);
  }
  uint8_t* d = (uint8_t*)dst;
  long long drs = (long long)dst_row_stride, dfs = (long long)dst_frame_stride;
  if (dst_mem_kind == DRFE_MEM_HOST) {
    if (!h->d_dst) DRFE_CUDA(cudaMalloc((void**)&h->d_dst, (size_t)h->max_batch * h->dw * h->dh * 4));
    d = h->d_dst; drs = (long long)dline; dfs = (long long)(dline * h->dh);
  }
  const dim3 grid((h->dw + 63) / 64, (h->dh + 3) / 4, nframes), block(64, 4);
#define DRFE_RS(T, CH) DRFE_LAUNCH((k_resize_linear<T, CH>), grid, block, 0, st, (const T*)s, srs, sfs, (T*)d, drs, dfs, h->dw, h->dh, h->d_xt, h->d_yt, h->sw == 2 * h->dw && h->sh == 2 * h->dh)
  if (esz == 1) { if (channels == 1) DRFE_RS(uint8_t, 1); else if (channels == 3) DRFE_RS(uint8_t, 3); else DRFE_RS(uint8_t, 4); }
  else if (esz == 2) DRFE_RS(uint16_t, 1);
  else DRFE_RS(float, 1);
#undef DRFE_RS
  if (dst_mem_kind == DRFE_MEM_HOST) {
    for (int f = 0; f < nframes; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync((uint8_t*)dst + (size_t)f * dst_frame_stride, dst_row_stride, h->d_dst + (size_t)f * dline * h->dh, dline, dline, h->dh, cudaMemcpyDeviceToHost, st));
    DRFE_CUDA(cudaStreamSynchronize(st));
  }
  return DRFE_OK;
}

}  // extern "C"
