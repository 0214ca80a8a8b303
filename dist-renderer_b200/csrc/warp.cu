// Two-view photometric warp (next-1 of SURVEY.md 8f): reprojection of view 1's hit points into view 2, depth-consistency
// test against view 2's rendered depth, bilinear colour sampling and the L1 colour loss -- one kernel forward, one backward.
//
// Replaces the dozen elementwise PyTorch ops of SDFRenderer_warp.get_valid_points / valid_points_depth / compute_loss_color
// (core/sdfrenderer/renderer_warp.py:18-101) and grid_sample_on_img (core/utils/loss_utils.py:9-25; the reference was
// written for torch 1.1, whose grid_sample convention is align_corners=True with zero padding: pixel coordinates are used
// as they are, taps outside the image contribute zero).
#include <cuda_runtime.h>
#include <math.h>
#include "common.cuh"

namespace dist {
namespace {

struct WarpCam {
  float Kinv[9], K[9];
  const float* R1;   // device [9]
  const float* c1;   // device [3]  camera centre of view 1 (-R1^T T1)
  const float* R2;   // device [9]
  const float* T2;   // device [3]
  int W, H;
};

// unit ray of pixel (x, y) of view 1, world frame (renderer.py:190-200)
__device__ __forceinline__ void warp_ray(const WarpCam& cam, const float* R, float x, float y, float (&ray)[3]) {
  float hc[3], v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) hc[i] = fmaf(cam.Kinv[i * 3 + 2], 1.f, fmaf(cam.Kinv[i * 3 + 1], y, cam.Kinv[i * 3] * x));
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = fmaf(R[6 + i], hc[2], fmaf(R[3 + i], hc[1], R[i] * hc[0]));
  const float nrm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-12f;
#pragma unroll
  for (int i = 0; i < 3; ++i) ray[i] = v[i] / nrm;
}

struct Taps { int x0, y0; float wx, wy; bool in00, in01, in10, in11; };
__device__ __forceinline__ Taps make_taps(float u, float v, int W, int H) {
  Taps t;
  const float fx = floorf(u), fy = floorf(v);
  t.x0 = (int)fx; t.y0 = (int)fy; t.wx = u - fx; t.wy = v - fy;
  const bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  t.in00 = xa && ya; t.in01 = xb && ya; t.in10 = xa && yb; t.in11 = xb && yb;
  return t;
}
// bilinear sample of channel c of an image stored [H][W][C] (zero padding)
__device__ __forceinline__ float sample(const float* img, int C, int c, const Taps& t, int W) {
  const float a = t.in00 ? img[((size_t)t.y0 * W + t.x0) * C + c] : 0.f, b = t.in01 ? img[((size_t)t.y0 * W + t.x0 + 1) * C + c] : 0.f;
  const float d = t.in10 ? img[((size_t)(t.y0 + 1) * W + t.x0) * C + c] : 0.f, e = t.in11 ? img[((size_t)(t.y0 + 1) * W + t.x0 + 1) * C + c] : 0.f;
  return (a * (1.f - t.wx) + b * t.wx) * (1.f - t.wy) + (d * (1.f - t.wx) + e * t.wx) * t.wy;
}
__device__ __forceinline__ bool finite_uv(float u, float v) { return fabsf(u) < 1e8f && fabsf(v) < 1e8f; }

// projection of view-1 pixel lp at depth z into view 2: xyz = K (R2 p + T2), p = c1 + ray z   (renderer_warp.py:22-32)
__device__ __forceinline__ void project(const WarpCam& cam, int lp, float z, float (&ray)[3], float (&p)[3], float (&xyz)[3]) {
  float R1[9], R2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { R1[i] = cam.R1[i]; R2[i] = cam.R2[i]; }
  warp_ray(cam, R1, (float)(lp % cam.W), (float)(lp / cam.W), ray);
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = ray[i] * z + cam.c1[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = fmaf(R2[i * 3 + 2], p[2], fmaf(R2[i * 3 + 1], p[1], R2[i * 3] * p[0])) + cam.T2[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) xyz[i] = fmaf(cam.K[i * 3 + 2], q[2], fmaf(cam.K[i * 3 + 1], q[1], cam.K[i * 3] * q[0]));
}

__global__ void k_warp_fwd(WarpCam cam, const float* Z1, const uint8_t* mask1, const float* depth2, const float* img1,
                           const float* img2, float thres, float* loss_sum, int32_t* count, uint8_t* keep, float* vis1,
                           float* vis2, int P) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  int n = 0;
  if (lp < P) {
    bool k = false;
    float c1v[3] = {0.f, 0.f, 0.f}, c2v[3] = {0.f, 0.f, 0.f};
    if (mask1[lp]) {
      float ray[3], p[3], xyz[3];
      project(cam, lp, Z1[lp], ray, p, xyz);
      const float u = xyz[0] / xyz[2], v = xyz[1] / xyz[2];
      if (finite_uv(u, v)) {
        const Taps t = make_taps(u, v, cam.W, cam.H);
        const float d2 = sample(depth2, 1, 0, t, cam.W);                     // renderer_warp.py:62-68
        const float e = xyz[2] - d2;
        if (e * e < thres) {                                                 // :70-71
          k = true;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            c1v[c] = img1[(size_t)lp * 3 + c];
            c2v[c] = sample(img2, 3, c, t, cam.W);                           // :81-83
            l += fabsf(c1v[c] - c2v[c]);                                     // :85
          }
          n = 1;
        }
      }
    }
    keep[lp] = k ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { vis1[(size_t)lp * 3 + c] = c1v[c]; vis2[(size_t)lp * 3 + c] = c2v[c]; }   // :96-99
  }
  for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if ((threadIdx.x & 31) == 0 && n) { atomicAdd(loss_sum, l); atomicAdd(count, n); }
}

// gscale[0] = dL / d(loss_sum).  Outputs: dZ1[P], d_ray1[3][P] (w.r.t. the unit rays of view 1), d_c1[3], dR2[9], dT2[3].
__global__ void k_warp_bwd(WarpCam cam, const float* Z1, const uint8_t* keep, const float* img1, const float* img2,
                           const float* gscale, float* dZ1, float* d_ray1, float* d_c1, float* dR2, float* dT2, int P) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.f;
  if (lp < P) {
    float dz = 0.f, dr[3] = {0.f, 0.f, 0.f};
    if (keep[lp]) {
      const float g = gscale[0];
      float ray[3], p[3], xyz[3];
      const float z = Z1[lp];
      project(cam, lp, z, ray, p, xyz);
      const float u = xyz[0] / xyz[2], v = xyz[1] / xyz[2];
      const Taps t = make_taps(u, v, cam.W, cam.H);
      float du = 0.f, dv = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = t.in00 ? img2[((size_t)t.y0 * cam.W + t.x0) * 3 + c] : 0.f, b = t.in01 ? img2[((size_t)t.y0 * cam.W + t.x0 + 1) * 3 + c] : 0.f;
        const float d = t.in10 ? img2[((size_t)(t.y0 + 1) * cam.W + t.x0) * 3 + c] : 0.f, e = t.in11 ? img2[((size_t)(t.y0 + 1) * cam.W + t.x0 + 1) * 3 + c] : 0.f;
        const float c2 = (a * (1.f - t.wx) + b * t.wx) * (1.f - t.wy) + (d * (1.f - t.wx) + e * t.wx) * t.wy;
        const float diff = img1[(size_t)lp * 3 + c] - c2;
        const float gc2 = -g * ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f));           // d|c1 - c2| / d c2
        du += gc2 * ((b - a) * (1.f - t.wy) + (e - d) * t.wy);
        dv += gc2 * ((d - a) * (1.f - t.wx) + (e - b) * t.wx);
      }
      // u = x / z3, v = y / z3
      const float iz = 1.f / xyz[2];
      const float dxyz[3] = {du * iz, dv * iz, -(du * xyz[0] + dv * xyz[1]) * iz * iz};
      float dq[3], dp[3], q[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) dq[i] = cam.K[i] * dxyz[0] + cam.K[3 + i] * dxyz[1] + cam.K[6 + i] * dxyz[2];   // K^T dxyz
#pragma unroll
      for (int i = 0; i < 3; ++i) dp[i] = cam.R2[i] * dq[0] + cam.R2[3 + i] * dq[1] + cam.R2[6 + i] * dq[2];       // R2^T dq
      (void)q;
      dz = ray[0] * dp[0] + ray[1] * dp[1] + ray[2] * dp[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) { dr[i] = z * dp[i]; acc[i] = dp[i]; acc[12 + i] = dq[i]; }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[3 + i * 3 + j] = dq[i] * p[j];                                           // dR2 = dq p^T
    }
    dZ1[lp] = dz;
#pragma unroll
    for (int i = 0; i < 3; ++i) d_ray1[(size_t)i * P + lp] = dr[i];
  }
#pragma unroll
  for (int i = 0; i < 15; ++i) {
    float t = acc[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t != 0.f) atomicAdd(i < 3 ? d_c1 + i : (i < 12 ? dR2 + (i - 3) : dT2 + (i - 12)), t);
  }
}

int make_warp_cam(const dist_camera_t* cam1, const float* K, const float* R2, const float* T2, WarpCam* w) {
  DIST_REQUIRE(cam1 && cam1->R && cam1->cam_pos && K && R2 && T2, "warp: null argument");
  DIST_REQUIRE(cam1->n_views <= 1 && cam1->row0 == 0 && cam1->n_rows == cam1->height && cam1->row_step == (cam1->row_group > 0 ? cam1->row_group : 1),
               "warp: view 1 must be one full image");
  for (int i = 0; i < 9; ++i) { w->Kinv[i] = cam1->Kinv[i]; w->K[i] = K[i]; }
  w->R1 = cam1->R; w->c1 = cam1->cam_pos; w->R2 = R2; w->T2 = T2; w->W = cam1->width; w->H = cam1->height;
  return DIST_OK;
}

}  // namespace
}  // namespace dist

using namespace dist;

extern "C" {

int dist_warp_loss_fwd(const dist_camera_t* cam1, const float* K_host, const float* R2, const float* T2, const float* Zdepth1,
                       const uint8_t* mask1, const float* depth2, const float* img1, const float* img2, float thres_depth,
                       float* loss_sum, int32_t* count, uint8_t* keep, float* vis1, float* vis2, void* stream) {
  WarpCam w;
  int rc = make_warp_cam(cam1, K_host, R2, T2, &w);
  if (rc) return rc;
  DIST_REQUIRE(Zdepth1 && mask1 && depth2 && img1 && img2 && loss_sum && count && keep && vis1 && vis2, "warp_loss_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = w.W * w.H;
  DIST_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
  DIST_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
  k_warp_fwd<<<(P + 255) / 256, 256, 0, st>>>(w, Zdepth1, mask1, depth2, img1, img2, thres_depth, loss_sum, count, keep, vis1, vis2, P);
  count_launch();
  DIST_CHECK_CUDA(cudaGetLastError());
  return DIST_OK;
}

int dist_warp_loss_bwd(const dist_camera_t* cam1, const float* K_host, const float* R2, const float* T2, const float* Zdepth1,
                       const uint8_t* keep, const float* img1, const float* img2, const float* gscale, float* dZdepth1,
                       float* d_ray1, float* d_cam_pos1, float* dR2, float* dT2, void* stream) {
  WarpCam w;
  int rc = make_warp_cam(cam1, K_host, R2, T2, &w);
  if (rc) return rc;
  DIST_REQUIRE(Zdepth1 && keep && img1 && img2 && gscale && dZdepth1 && d_ray1 && d_cam_pos1 && dR2 && dT2, "warp_loss_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = w.W * w.H;
  DIST_CHECK_CUDA(cudaMemsetAsync(d_cam_pos1, 0, 3 * sizeof(float), st));
  DIST_CHECK_CUDA(cudaMemsetAsync(dR2, 0, 9 * sizeof(float), st));
  DIST_CHECK_CUDA(cudaMemsetAsync(dT2, 0, 3 * sizeof(float), st));
  k_warp_bwd<<<(P + 255) / 256, 256, 0, st>>>(w, Zdepth1, keep, img1, img2, gscale, dZdepth1, d_ray1, d_cam_pos1, dR2, dT2, P);
  count_launch();
  DIST_CHECK_CUDA(cudaGetLastError());
  return DIST_OK;
}

}  // extern "C"
