// Caller-side step pieces:
//   * cross-entropy + top-k accuracy + running meters in two launches, nothing read back
//     (train.py:493/614 criterion, :518 accuracy, utils/utils.py:72-85, meters :520-524)
//   * multi-tensor Adam / SGD-momentum update with the data-parallel 1/world gradient scale folded in
//     (train.py:319-336 optimizers, step at :529/:651)
#include "common.cuh"

namespace bdbnn {

__device__ __forceinline__ float blk_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = lane < nw ? sh[lane] : -INFINITY;
  return warp_max(t);
}
__device__ __forceinline__ float blk_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = lane < nw ? sh[lane] : 0.f;
  return warp_sum(t);
}

// One block per sample: row_loss = logsumexp(z) - z[target]; grad = (softmax(z) - onehot)/N_valid;
// row_rank = number of classes that beat the target (ties: lower index wins, like a stable top-k).
// nn.CrossEntropyLoss() edge cases (train.py:298 constructs it with the defaults): a target equal to
// ignore_index (-100) contributes no loss and no gradient and is left out of the mean's divisor
// (N_valid = number of non-ignored rows; all ignored -> NaN like torch); any other target outside
// [0, C) is a caller bug — torch raises a device-side assert, this kernel traps (sticky launch failure).
constexpr int64_t kIgnoreIndex = -100;
__global__ void __launch_bounds__(128)
ce_row_kernel(const float* __restrict__ z, const int64_t* __restrict__ target, int32_t N, int32_t C,
              float* __restrict__ row_loss, int32_t* __restrict__ row_rank, float* __restrict__ grad) {
  __shared__ float sh[32];
  const int n = blockIdx.x;
  const float* zr = z + int64_t(n) * C;
  const int64_t t = target[n];
  const bool t_ok = t >= 0 && t < C;
  if (!t_ok && t != kIgnoreIndex) __trap();
  const float zt = t_ok ? zr[t] : -INFINITY;
  float valid = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) valid += target[i] != kIgnoreIndex ? 1.f : 0.f;
  valid = blk_sum(valid, sh);
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, zr[c]);
  m = blk_max(m, sh);
  float e = 0.f, beat = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = zr[c];
    e += expf(v - m);
    beat += (v > zt || (v == zt && c < t)) ? 1.f : 0.f;
  }
  e = blk_sum(e, sh);
  beat = blk_sum(beat, sh);
  const float lse = m + logf(e);
  if (grad != nullptr) {
    const float inv_e = 1.0f / e, invN = 1.0f / valid;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      grad[int64_t(n) * C + c] = t_ok ? (expf(zr[c] - m) * inv_e - (c == t ? 1.f : 0.f)) * invN : 0.f;
  }
  if (threadIdx.x == 0) {
    row_loss[n] = t_ok ? lse - zt : 0.f;
    row_rank[n] = t_ok ? int32_t(beat) : C;
  }
}

// loss = mean of the rows (fixed order), acc[k] = 100 * #(rank < topk[k]) / N; optional running meters
// meters = {sum loss*N, sum acc1*N, sum acc5*N, samples} (AverageMeter.update(val, n), utils/utils.py).
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float* __restrict__ row_loss, const int32_t* __restrict__ row_rank,
                   const int64_t* __restrict__ target, int32_t N, int32_t k1,
                   int32_t k2, float* __restrict__ loss_out, float* __restrict__ acc_out,
                   double* __restrict__ meters) {
  __shared__ double red[4][8];
  double a = 0.0, c1 = 0.0, c2 = 0.0, nv = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    a += double(row_loss[i]);
    const int r = row_rank[i];
    c1 += r < k1 ? 1.0 : 0.0;
    c2 += r < k2 ? 1.0 : 0.0;
    nv += target[i] != kIgnoreIndex ? 1.0 : 0.0;
  }
  a = warp_sum(a); c1 = warp_sum(c1); c2 = warp_sum(c2); nv = warp_sum(nv);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = c1; red[2][threadIdx.x >> 5] = c2;
    red[3][threadIdx.x >> 5] = nv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, t1 = 0.0, t2 = 0.0, tv = 0.0;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) { ta += red[0][i]; t1 += red[1][i]; t2 += red[2][i]; tv += red[3][i]; }
    const float loss = float(ta / tv);           // mean over the non-ignored rows (0/0 = NaN like torch)
    const float acc1 = float(t1 * 100.0 / double(N)), acc2 = float(t2 * 100.0 / double(N));
    loss_out[0] = loss;
    acc_out[0] = acc1;
    acc_out[1] = acc2;
    if (meters != nullptr) {
      meters[0] += double(loss) * N;
      meters[1] += double(acc1) * N;
      meters[2] += double(acc2) * N;
      meters[3] += double(N);
    }
  }
}

// ---- multi-tensor optimizers ---------------------------------------------------------------------
constexpr int kOptMaxTensors = 48;
struct OptTable {
  float* p[kOptMaxTensors];
  const float* g[kOptMaxTensors];
  float* m[kOptMaxTensors];
  float* v[kOptMaxTensors];
  int64_t n[kOptMaxTensors];
  float wd[kOptMaxTensors];
  float lr[kOptMaxTensors];
};

// torch.optim.Adam (L2 weight decay added to the gradient, no amsgrad), blockIdx.y = tensor.
// step_dev != NULL (CUDA-graph mode): the 1-based step count and the learning rates are read from device memory
// (step_dev[0]; lr_dev[lr_off + t]), so a captured launch stays valid while both advance between replays.
__global__ void __launch_bounds__(256)
adam_multi_kernel(const OptTable tab, float beta1, float beta2, float eps, float bc1, float rsqrt_bc2,
                  float grad_scale, const float* __restrict__ step_dev, const float* __restrict__ lr_dev, int lr_off) {
  const int t = blockIdx.y;
  float* __restrict__ p = tab.p[t];
  const float* __restrict__ g = tab.g[t];
  float* __restrict__ m = tab.m[t];
  float* __restrict__ v = tab.v[t];
  const int64_t n = tab.n[t];
  float lr = tab.lr[t];
  if (step_dev != nullptr) {
    const float step = __ldg(step_dev);
    bc1 = 1.0f - powf(beta1, step);
    rsqrt_bc2 = rsqrtf(1.0f - powf(beta2, step));
    if (lr_dev != nullptr) lr = __ldg(lr_dev + lr_off + t);
  }
  const float wd = tab.wd[t], step_size = lr / bc1;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * grad_scale);
    const float mv = fmaf(beta1, m[i], (1.0f - beta1) * gv);
    const float vv = fmaf(beta2, v[i], (1.0f - beta2) * gv * gv);
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) * rsqrt_bc2 + eps;
    p[i] = pv - step_size * (mv / denom);
  }
}

// torch.optim.SGD with momentum (dampening 0, no nesterov): buf = g' on the first step, else mu*buf + g'.
__global__ void __launch_bounds__(256)
sgd_multi_kernel(const OptTable tab, float momentum, int first_step, float grad_scale,
                 const float* __restrict__ lr_dev, int lr_off) {
  const int t = blockIdx.y;
  float* __restrict__ p = tab.p[t];
  const float* __restrict__ g = tab.g[t];
  float* __restrict__ m = tab.m[t];
  const int64_t n = tab.n[t];
  const float wd = tab.wd[t], lr = lr_dev != nullptr ? __ldg(lr_dev + lr_off + t) : tab.lr[t];
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * grad_scale);
    float d = gv;
    if (m != nullptr) {
      d = first_step ? gv : fmaf(momentum, m[i], gv);
      m[i] = d;
    }
    p[i] = pv - lr * d;
  }
}

__global__ void step_inc_kernel(float* step) { step[0] += 1.0f; }

static unsigned opt_blocks(const int64_t* n, int count) {
  int64_t mx = 1;
  for (int i = 0; i < count; ++i) mx = n[i] > mx ? n[i] : mx;
  int64_t b = (mx + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = int64_t(num_sms()) * 2;
  if (b > cap) b = cap;
  return unsigned(b < 1 ? 1 : b);
}

}  // namespace bdbnn

using namespace bdbnn;

extern "C" int bdbnn_ce_topk_fwd_bwd(const float* logits, const int64_t* target, int32_t N, int32_t C, int32_t k1,
                                     int32_t k2, float* row_loss_ws, int32_t* row_rank_ws, float* loss_out,
                                     float* acc_out, float* grad_logits, double* meters, void* stream) {
  BDBNN_REQUIRE(N > 0 && C > 0 && k1 > 0 && k2 > 0, "ce_topk: bad N/C/k");
  BDBNN_REQUIRE(logits && target && row_loss_ws && row_rank_ws && loss_out && acc_out, "ce_topk: NULL pointer");
  cudaStream_t st = cudaStream_t(stream);
  ce_row_kernel<<<N, 128, 0, st>>>(logits, target, N, C, row_loss_ws, row_rank_ws, grad_logits);
  int rc = check_launch("ce_row_kernel");
  if (rc) return rc;
  ce_finalize_kernel<<<1, 256, 0, st>>>(row_loss_ws, row_rank_ws, target, N, k1, k2, loss_out, acc_out, meters);
  return check_launch("ce_finalize_kernel");
}

static int fill_table(OptTable& tab, float* const* p, const float* const* g, float* const* m, float* const* v,
                      const int64_t* n, const float* wd, const float* lr, int off, int cnt, bool need_v,
                      bool need_m) {
  for (int i = 0; i < cnt; ++i) {
    BDBNN_REQUIRE(p[off + i] && g[off + i] && n[off + i] >= 0, "optimizer: NULL parameter/gradient %d", off + i);
    BDBNN_REQUIRE(!need_m || m[off + i], "optimizer: NULL first-moment buffer %d", off + i);
    BDBNN_REQUIRE(!need_v || v[off + i], "optimizer: NULL second-moment buffer %d", off + i);
    tab.p[i] = p[off + i]; tab.g[i] = g[off + i];
    tab.m[i] = m ? m[off + i] : nullptr;
    tab.v[i] = v ? v[off + i] : nullptr;
    tab.n[i] = n[off + i]; tab.wd[i] = wd[off + i]; tab.lr[i] = lr[off + i];
  }
  return BDBNN_OK;
}

static int adam_launch(float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                       float* const* exp_avg_sq_host, const int64_t* numel_host, const float* weight_decay_host,
                       const float* lr_host, int32_t count, float beta1, float beta2, float eps, int64_t step,
                       const float* step_dev, const float* lr_dev, float grad_scale, void* stream) {
  BDBNN_REQUIRE(count == 0 || (params_host && grads_host && exp_avg_host && exp_avg_sq_host && numel_host &&
                               weight_decay_host && lr_host), "optim_adam: NULL table");
  const float bc1 = step_dev ? 1.0f : float(1.0 - pow(double(beta1), double(step)));
  const float rsqrt_bc2 = step_dev ? 1.0f : float(1.0 / sqrt(1.0 - pow(double(beta2), double(step))));
  for (int off = 0; off < count; off += kOptMaxTensors) {
    const int cnt = count - off < kOptMaxTensors ? count - off : kOptMaxTensors;
    OptTable tab;
    memset(&tab, 0, sizeof(tab));
    int rc = fill_table(tab, params_host, grads_host, exp_avg_host, exp_avg_sq_host, numel_host, weight_decay_host,
                        lr_host, off, cnt, true, true);
    if (rc) return rc;
    dim3 grid(opt_blocks(numel_host + off, cnt), unsigned(cnt));
    adam_multi_kernel<<<grid, 256, 0, cudaStream_t(stream)>>>(tab, beta1, beta2, eps, bc1, rsqrt_bc2, grad_scale,
                                                              step_dev, lr_dev, off);
    rc = check_launch("adam_multi_kernel");
    if (rc) return rc;
  }
  return BDBNN_OK;
}

extern "C" int bdbnn_optim_adam_multi(float* const* params_host, const float* const* grads_host,
                                      float* const* exp_avg_host, float* const* exp_avg_sq_host,
                                      const int64_t* numel_host, const float* weight_decay_host,
                                      const float* lr_host, int32_t count, float beta1, float beta2, float eps,
                                      int64_t step, float grad_scale, void* stream) {
  BDBNN_REQUIRE(count >= 0 && step >= 1, "optim_adam: bad count/step");
  return adam_launch(params_host, grads_host, exp_avg_host, exp_avg_sq_host, numel_host, weight_decay_host, lr_host,
                     count, beta1, beta2, eps, step, nullptr, nullptr, grad_scale, stream);
}

extern "C" int bdbnn_optim_step_inc(float* step_dev, void* stream) {
  BDBNN_REQUIRE(step_dev != nullptr, "optim_step_inc: NULL step");
  step_inc_kernel<<<1, 1, 0, cudaStream_t(stream)>>>(step_dev);
  return check_launch("step_inc_kernel");
}

extern "C" int bdbnn_optim_adam_multi_graph(float* const* params_host, const float* const* grads_host,
                                            float* const* exp_avg_host, float* const* exp_avg_sq_host,
                                            const int64_t* numel_host, const float* weight_decay_host,
                                            int32_t count, float beta1, float beta2, float eps,
                                            const float* step_dev, const float* lr_dev, float grad_scale,
                                            void* stream) {
  BDBNN_REQUIRE(count >= 0 && step_dev && lr_dev, "optim_adam_graph: bad count / NULL device step or lr");
  float lr0[kOptMaxTensors] = {0};
  // learning rates come from lr_dev; the host table is only a placeholder of the right length
  const float* lr_host = nullptr;
  float* lr_tmp = nullptr;
  if (count > kOptMaxTensors) {
    lr_tmp = static_cast<float*>(calloc(size_t(count), sizeof(float)));
    BDBNN_REQUIRE(lr_tmp != nullptr, "optim_adam_graph: out of host memory");
    lr_host = lr_tmp;
  } else {
    lr_host = lr0;
  }
  const int rc = adam_launch(params_host, grads_host, exp_avg_host, exp_avg_sq_host, numel_host, weight_decay_host,
                             lr_host, count, beta1, beta2, eps, 1, step_dev, lr_dev, grad_scale, stream);
  free(lr_tmp);
  return rc;
}

static int sgd_launch(float* const* params_host, const float* const* grads_host, float* const* momentum_buf_host,
                      const int64_t* numel_host, const float* weight_decay_host, const float* lr_host, int32_t count,
                      float momentum, int32_t first_step, float grad_scale, const float* lr_dev, void* stream) {
  BDBNN_REQUIRE(count >= 0, "optim_sgd: bad count");
  BDBNN_REQUIRE(count == 0 || (params_host && grads_host && numel_host && weight_decay_host && lr_host),
                "optim_sgd: NULL table");
  BDBNN_REQUIRE(momentum == 0.0f || momentum_buf_host, "optim_sgd: momentum needs buffers");
  for (int off = 0; off < count; off += kOptMaxTensors) {
    const int cnt = count - off < kOptMaxTensors ? count - off : kOptMaxTensors;
    OptTable tab;
    memset(&tab, 0, sizeof(tab));
    int rc = fill_table(tab, params_host, grads_host, momentum != 0.0f ? momentum_buf_host : nullptr, nullptr,
                        numel_host, weight_decay_host, lr_host, off, cnt, false, momentum != 0.0f);
    if (rc) return rc;
    dim3 grid(opt_blocks(numel_host + off, cnt), unsigned(cnt));
    sgd_multi_kernel<<<grid, 256, 0, cudaStream_t(stream)>>>(tab, momentum, first_step, grad_scale, lr_dev, off);
    rc = check_launch("sgd_multi_kernel");
    if (rc) return rc;
  }
  return BDBNN_OK;
}

extern "C" int bdbnn_optim_sgd_multi(float* const* params_host, const float* const* grads_host,
                                     float* const* momentum_buf_host, const int64_t* numel_host,
                                     const float* weight_decay_host, const float* lr_host, int32_t count,
                                     float momentum, int32_t first_step, float grad_scale, void* stream) {
  return sgd_launch(params_host, grads_host, momentum_buf_host, numel_host, weight_decay_host, lr_host, count,
                    momentum, first_step, grad_scale, nullptr, stream);
}

extern "C" int bdbnn_optim_sgd_multi_graph(float* const* params_host, const float* const* grads_host,
                                           float* const* momentum_buf_host, const int64_t* numel_host,
                                           const float* weight_decay_host, int32_t count, float momentum,
                                           const float* lr_dev, float grad_scale, void* stream) {
  BDBNN_REQUIRE(count >= 0 && lr_dev, "optim_sgd_graph: bad count / NULL device lr");
  float* lr_tmp = static_cast<float*>(calloc(size_t(count > 0 ? count : 1), sizeof(float)));
  BDBNN_REQUIRE(lr_tmp != nullptr, "optim_sgd_graph: out of host memory");
  // zero-initialised momentum buffers make the first step identical to torch's (buf = mu*0 + g)
  const int rc = sgd_launch(params_host, grads_host, momentum_buf_host, numel_host, weight_decay_host, lr_tmp, count,
                            momentum, 0, grad_scale, lr_dev, stream);
  free(lr_tmp);
  return rc;
}
