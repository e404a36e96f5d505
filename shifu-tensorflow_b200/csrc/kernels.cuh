// HBM-bound kernels of the step: mini-batch load/cast/transpose (K1), output layer + loss + its
// backward (K3/K4), fused multi-tensor optimizer (K7), gradient accumulation.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

// Where the current mini-batch lives (HBM-resident training set or the H2D staging area).
// Written by set_batch_kernel right before the captured step graph runs, so the graph itself
// never changes (res/ssgd_monitor.py:272-276 builds the same feed_dict from x_batch[i]).
struct BatchDesc {
  const float* X;  // [rows, F] row-major fp32
  const float* y;  // [rows]
  const float* w;  // [rows]
  float lr_t;      // Adam: lr * sqrt(1-b2^t)/(1-b1^t); others: lr
  float gscale;    // 1/world (and 1/n_accumulated for the epoch-sync schedule)
};

static __global__ void set_batch_kernel(BatchDesc* d, const float* X, const float* y, const float* w, float lr_t, float gscale) {
  d->X = X; d->y = y; d->w = w; d->lr_t = lr_t; d->gscale = gscale;
}

// step scalars (device): [0] = sum_i w_i * per-row loss, [1] = n_nz (count of non-zero weights)
enum { SCAL_LOSS_SUM = 0, SCAL_NNZ = 1, SCAL_COUNT = 4 };

// ------------------------------------------------------------------------------------------------
// K1 mini-batch load.  fp32 rows -> (bf16 mode) row-major bf16 [rows, ldF] AND transposed bf16
// [F, ldB] (the K-major operand of the layer-0 dW GEMM); (fp32 mode) fp32 copy into the batch buffer.
// 32x32 tiles through shared memory so both the read and both writes are coalesced.
// Block (0,0) additionally counts the non-zero sample weights of the batch (n_nz of
// SUM_BY_NONZERO_WEIGHTS, res/ssgd_monitor.py:129).
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256)
load_batch_kernel(const BatchDesc* __restrict__ desc, int rows, int F, __nv_bfloat16* __restrict__ Xb, int ldF,
                  __nv_bfloat16* __restrict__ XbT, int ldB, float* __restrict__ Xf, float* __restrict__ scal) {
  __shared__ float tile[32][33];
  const float* __restrict__ X = desc->X;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < F) v = __ldg(X + static_cast<size_t>(r) * F + c);
    if constexpr (BF16) {
      tile[ty + 8 * i][tx] = v;
      if (r < rows && c < ldF) Xb[static_cast<size_t>(r) * ldF + c] = __float2bfloat16_rn(v);
    } else {
      if (r < rows && c < F) Xf[static_cast<size_t>(r) * F + c] = v;
    }
  }
  if constexpr (BF16) {
    if (XbT != nullptr) {  // the scorer needs no transposed copy (no dW GEMM)
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < F && r < rows) XbT[static_cast<size_t>(c) * ldB + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
      }
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    const float* __restrict__ w = desc->w;
    float cnt = 0.f;
    for (int i = threadIdx.x; i < rows; i += 256) cnt += (__ldg(w + i) != 0.f) ? 1.f : 0.f;
    cnt = warp_sum(cnt);
    __shared__ float part[8];
    if (tx == 0) part[ty] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += part[i];
      scal[SCAL_NNZ] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3+K4 (+ output-layer backward): y_hat = sigmoid(A_L w_o + b_o); loss = sum w (y_hat-y)^2 / n_nz
// (res/ssgd_monitor.py:121,129) or the sigmoid-CE variant; d z_hat; then the rank-1 backward
//   dZ_L[r,j] = dz_r * w_o[j] * act'(A_L[r,j])     (row-major + transposed bf16 copies)
//   dw_o[j] += sum_r dz_r A_L[r,j],  db_o += sum_r dz_r,  db_L[j] += sum_r dZ_L[r,j]
// out = 1 is GEMV-class: CUDA cores, one pass over A_L.  Each block owns 32 rows.
// ------------------------------------------------------------------------------------------------
struct OutLayerParams {
  int rows, H, ldA;          // A_L is [rows, ldA] with H valid columns
  const void* A;             // bf16 or fp32
  const float* wo;           // [H]
  const float* bo;           // [1]
  const BatchDesc* desc;     // y, w
  float* scal;               // SCAL_*
  int loss, act;             // sb_loss, activation of hidden layer L
  int do_bwd;                // 0: forward (+loss) only
  int do_loss;               // 0: scores only (no y / w access)
  float* yhat;               // nullable [rows]
  void* dZ; int ld_dZ;       // [rows, ld_dZ] bf16 or fp32
  __nv_bfloat16* dZT; int ld_dZT;  // [H, ld_dZT] (bf16 mode only)
  float* g_wo; float* g_bo; float* g_bL;  // gradient slots (atomic accumulate)
};

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
out_layer_kernel(const OutLayerParams p) {
  __shared__ float dz_row[32];
  __shared__ float tileT[128][33];  // [column in chunk][row] for the transposed write
  __shared__ float blk_red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * 32;
  const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
  const float bo = __ldg(p.bo);
  const float nnz = p.do_loss ? p.scal[SCAL_NNZ] : 1.f;
  const float inv_nnz = nnz > 0.f ? 1.f / nnz : 0.f;

  // ---- phase 1: one warp per row (4 rows per warp) ----
  float loss_part = 0.f;
  for (int i = 0; i < 4; ++i) {
    const int rl = warp * 4 + i, r = r0 + rl;
    float z = 0.f;
    if (r < p.rows) {
      const T* ar = A + static_cast<size_t>(r) * p.ldA;
      for (int j = lane; j < p.H; j += 32) z = fmaf(ld_as_float<T>(ar + j), __ldg(p.wo + j), z);
    }
    z = warp_sum(z) + bo;
    if (lane == 0) {
      float dz = 0.f;
      if (r < p.rows) {
        const float yh = sigmoidf_stable(z);
        if (p.yhat) p.yhat[r] = yh;
        if (p.do_loss) {
          const float y = __ldg(p.desc->y + r), w = __ldg(p.desc->w + r);
          if (p.loss == SB_LOSS_MSE) {
            const float d = yh - y;
            loss_part += w * d * d;
            dz = 2.f * w * d * yh * (1.f - yh) * inv_nnz;
          } else {
            loss_part += w * (fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z))));
            dz = w * (yh - y) * inv_nnz;
          }
        }
      }
      dz_row[rl] = dz;
    }
  }
  if (p.do_loss) {
    if (lane == 0) blk_red[warp] = loss_part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += blk_red[i];
      atomicAdd(p.scal + SCAL_LOSS_SUM, s);
    }
  }
  if (!p.do_bwd) return;
  __syncthreads();

  // ---- phase 2: rank-1 backward over column chunks of 128 ----
  float dbo = 0.f;
  if (tid < 32) {
    dbo = warp_sum(dz_row[tid]);
    if (tid == 0) atomicAdd(p.g_bo, dbo);
  }
  T* __restrict__ dZ = reinterpret_cast<T*>(p.dZ);
  for (int c0 = 0; c0 < p.H; c0 += 128) {
    const int cl = tid & 127, j = c0 + cl;  // column handled by this thread
    const int rh = tid >> 7;                // row half: rows rh*16 .. rh*16+15
    float s_dw = 0.f, s_db = 0.f;
    const float woj = (j < p.H) ? __ldg(p.wo + j) : 0.f;
    for (int i = 0; i < 16; ++i) {
      const int rl = rh * 16 + i, r = r0 + rl;
      float g = 0.f;
      if (r < p.rows && j < p.H) {
        const float a = ld_as_float<T>(A + static_cast<size_t>(r) * p.ldA + j);
        const float dz = dz_row[rl];
        g = dz * woj * act_grad_from_out(a, p.act);
        s_dw = fmaf(dz, a, s_dw);
        s_db += g;
        st_from_float<T>(dZ + static_cast<size_t>(r) * p.ld_dZ + j, g);
      }
      if (p.dZT) tileT[cl][rl] = g;
    }
    if (j < p.H) {
      atomicAdd(p.g_wo + j, s_dw);
      atomicAdd(p.g_bL + j, s_db);
    }
    if (p.dZT) {
      __syncthreads();
      // transposed write: 32 consecutive rows of one column = 64 contiguous bytes
      for (int q = warp; q < 128; q += 8) {
        const int jj = c0 + q, r = r0 + lane;
        if (jj < p.H && r < p.rows) p.dZT[static_cast<size_t>(jj) * p.ld_dZT + r] = __float2bfloat16_rn(tileT[q][lane]);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K7 fused multi-tensor optimizer over the flat parameter vector (TF 1.x kernel forms: ApplyAdadelta
// res/ssgd_monitor.py:138, ApplyAdam res/ssgd.py:57, ApplyGradientDescent res/ssgd_monitor_bk.py:81,
// ApplyMomentum).  Reads the (all-reduced) gradient once, updates fp32 master weights + state, and in
// bf16 mode refreshes the two bf16 shadow copies the GEMMs consume: W^T [out, ld_in] (forward B operand)
// and W [in, ld_out] (dA B operand), transposing 32x32 tiles through shared memory.
// ------------------------------------------------------------------------------------------------
struct OptHyper {
  int kind;
  float rho, eps, beta1, beta2, momentum;
};

__device__ __forceinline__ float opt_update(const OptHyper& h, float lr_t, float theta, float g, float& s1, float& s2) {
  switch (h.kind) {
    case SB_OPT_SGD:
      return theta - lr_t * g;
    case SB_OPT_MOMENTUM:
      s1 = s1 * h.momentum + g;
      return theta - lr_t * s1;
    case SB_OPT_ADAM:
      s1 = s1 + (g - s1) * (1.f - h.beta1);
      s2 = s2 + (g * g - s2) * (1.f - h.beta2);
      return theta - lr_t * s1 / (sqrtf(s2) + h.eps);
    default: {  // SB_OPT_ADADELTA
      s1 = s1 * h.rho + g * g * (1.f - h.rho);
      const float upd = sqrtf(s2 + h.eps) / sqrtf(s1 + h.eps) * g;
      s2 = s2 * h.rho + upd * upd * (1.f - h.rho);
      return theta - upd * lr_t;
    }
  }
}

// One work item per block: either a 32x32 tile of a hidden-layer weight matrix (with shadows) or a
// 1024-element run of "plain" parameters (biases, output layer, or everything in fp32 mode).
struct OptWork {
  long long off;        // flat offset of the matrix / run
  int kind;             // 0 = plain run, 1 = weight tile
  int count;            // plain: elements in this run (<= 1024)
  int in_dim, out_dim;  // tile: matrix dims
  int ti, to;           // tile: tile row (in) / col (out) index
  __nv_bfloat16* Wt; int ld_in;   // [out, ld_in]
  __nv_bfloat16* Wn; int ld_out;  // [in, ld_out]
};

static __global__ void __launch_bounds__(256)
optimizer_kernel(const OptWork* __restrict__ work, const BatchDesc* __restrict__ desc, OptHyper h,
                 float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ s1, float* __restrict__ s2) {
  __shared__ float tile[32][33];
  const OptWork wk = work[blockIdx.x];
  const float lr_t = desc->lr_t, gs = desc->gscale;
  if (wk.kind == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i;
      if (e < wk.count) {
        const long long idx = wk.off + e;
        float a = s1[idx], b = s2[idx];
        const float t = opt_update(h, lr_t, theta[idx], grad[idx] * gs, a, b);
        theta[idx] = t; s1[idx] = a; s2[idx] = b;
      }
    }
    return;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = wk.ti * 32, o0 = wk.to * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, o = o0 + tx;
    float t = 0.f;
    if (i < wk.in_dim && o < wk.out_dim) {
      const long long idx = wk.off + static_cast<long long>(i) * wk.out_dim + o;
      float a = s1[idx], b = s2[idx];
      t = opt_update(h, lr_t, theta[idx], grad[idx] * gs, a, b);
      theta[idx] = t; s1[idx] = a; s2[idx] = b;
      if (wk.Wn) wk.Wn[static_cast<size_t>(i) * wk.ld_out + o] = __float2bfloat16_rn(t);
    }
    tile[ty + 8 * k][tx] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = o0 + ty + 8 * k, i = i0 + tx;
    if (i < wk.in_dim && o < wk.out_dim) wk.Wt[static_cast<size_t>(o) * wk.ld_in + i] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

// Refresh the bf16 shadows from the fp32 master without touching state (after set_params / restore).
static __global__ void __launch_bounds__(256)
shadow_refresh_kernel(const OptWork* __restrict__ work, const float* __restrict__ theta) {
  __shared__ float tile[32][33];
  const OptWork wk = work[blockIdx.x];
  if (wk.kind == 0) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = wk.ti * 32, o0 = wk.to * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, o = o0 + tx;
    float t = 0.f;
    if (i < wk.in_dim && o < wk.out_dim) {
      t = theta[wk.off + static_cast<long long>(i) * wk.out_dim + o];
      if (wk.Wn) wk.Wn[static_cast<size_t>(i) * wk.ld_out + o] = __float2bfloat16_rn(t);
    }
    tile[ty + 8 * k][tx] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = o0 + ty + 8 * k, i = i0 + tx;
    if (i < wk.in_dim && o < wk.out_dim) wk.Wt[static_cast<size_t>(o) * wk.ld_in + i] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

// acc += g  (epoch-sync schedule: ConditionalAccumulator.apply_grad, res/ssgd_monitor.py:136-141)
static __global__ void axpy_kernel(float* __restrict__ acc, const float* __restrict__ g, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) acc[i] += g[i];
}
static __global__ void scale_kernel(float* __restrict__ g, const BatchDesc* __restrict__ desc, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) g[i] *= desc->gscale;
}
static __global__ void fill_kernel(float* __restrict__ p, float v, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
// f32 -> bf16 with arbitrary leading dims (test hook operand staging)
static __global__ void cast_bf16_kernel(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* __restrict__ dst, int ld) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < static_cast<long long>(rows) * cols) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    dst[static_cast<size_t>(r) * ld + c] = __float2bfloat16_rn(src[i]);
  }
}

}  // namespace sb
