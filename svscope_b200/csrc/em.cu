// Fused E/M-step kernel of the categorical sequence mixture model (svs_em_batch).
//
// Replaces ReadsCluster.EM (src/ReadsCluster.py:190-209): pitheta_updating (:162-188, M),
// gamma_updating (:132-155, E) and the log-likelihood of the LAST iteration (:104-122; the
// reference evaluates it every iteration but BIC only reads likelihood[-1], :216).
// One CTA per (window, K) task, FP64 throughout.  Per iteration the feature columns are
// processed in tiles: the tile of X is staged in shared memory, one thread per column
// accumulates the weighted symbol counts (M) and leaves log(clip(theta)) in shared memory,
// then the threads switch to a (read, part) layout and add the gathered log-probabilities
// into per-read sums (E).  theta is written to global memory only in the last iteration.
// No tensor cores: the E-step "dot" is a gather-add over a one-hot, not a dense contraction.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "context.h"

namespace svs {
namespace {

constexpr double kEps = 1e-10;  // CheckParam clamp (src/ReadsCluster.py:70-74)
constexpr int kA = 5;           // alphabet A,T,C,G,-

struct EmTask {
  const int8_t* X;         // N x nf (device)
  const int32_t* labels;   // N hard labels 0..K-1, or nullptr -> start from theta/pi
  double* gamma;           // N x K
  double* theta;           // K x nf x 5 (in when labels == nullptr, always out)
  double* pi;              // K
  double* loglik;          // N
  int32_t* status;         // -1 done, else index of the M-step that hit the re-draw condition
  int32_t N, nf, n_steps, ft;
};

__device__ __forceinline__ double clip01(double x) { return fmin(fmax(x, kEps), 1.0 - kEps); }

template <int K, int T>
__global__ void __launch_bounds__(T) em_kernel(const EmTask* __restrict__ tasks, const int32_t* __restrict__ ids) {
  const EmTask t = tasks[ids[blockIdx.x]];
  const int N = t.N, nf = t.nf, FT = t.ft;
  const int tid = threadIdx.x;
  extern __shared__ __align__(16) unsigned char em_smem[];
  double* gam = reinterpret_cast<double*>(em_smem);          // N*K
  double* logt = gam + static_cast<size_t>(N) * K;            // FT*K*5
  double* part = logt + static_cast<size_t>(FT) * K * kA;     // T*K partial sums / scratch
  double* colsum = part + static_cast<size_t>(T) * K;         // K
  double* pis = colsum + K;                                    // K
  int8_t* xt = reinterpret_cast<int8_t*>(pis + K);            // FT*N tile of X, [f][n]
  __shared__ int bad_flag;

  const int parts = max(1, T / N);          // threads per read in the E layout (N <= T)
  const int en = tid % N, ep = tid / N;     // read / part of this thread
  const bool e_active = tid < parts * N;

  const bool from_theta = (t.labels == nullptr);
  if (!from_theta) {
    for (int idx = tid; idx < N * K; idx += T) gam[idx] = (t.labels[idx / K] == (idx % K)) ? 1.0 : 0.0;
  } else if (tid < K) {
    pis[tid] = t.pi[tid];
  }
  if (tid == 0) bad_flag = 0;
  __syncthreads();

  double S[K];  // per-read sum over features of log theta (valid in part 0 after reduction)
  for (int it = 0; it <= t.n_steps; ++it) {
    const bool do_m = !(it == 0 && from_theta);
    const bool last = (it == t.n_steps);
    if (do_m) {
      if (tid < K) {
        double sum = 0.0;
        for (int n = 0; n < N; ++n) sum += gam[n * K + tid];
        colsum[tid] = sum;
        const double p = sum / static_cast<double>(N);
        pis[tid] = p;
        if (p * static_cast<double>(N) < 1.0 || p != p) bad_flag = 1;
      }
      __syncthreads();
      if (bad_flag) {
        if (tid == 0) *t.status = it;
        return;
      }
    }
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;

    for (int f0 = 0; f0 < nf; f0 += FT) {
      const int fn = min(FT, nf - f0);
      // stage the tile of X as [f][n]
      for (int idx = tid; idx < fn * N; idx += T) {
        const int n = idx / fn, f = idx % fn;
        xt[f * N + n] = t.X[static_cast<size_t>(n) * nf + f0 + f];
      }
      __syncthreads();
      for (int f = tid; f < fn; f += T) {
        double* lt = logt + static_cast<size_t>(f) * K * kA;
        if (do_m) {
          double cnt[K][kA];
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int a = 0; a < kA; ++a) cnt[k][a] = 0.0;
          for (int n = 0; n < N; ++n) {
            const int xv = xt[f * N + n];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double g = gam[n * K + k];
#pragma unroll
              for (int a = 0; a < kA; ++a) cnt[k][a] += (xv == a) ? g : 0.0;
            }
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int a = 0; a < kA; ++a) {
              const double th = cnt[k][a] / colsum[k];
              if (last && t.theta) t.theta[(static_cast<size_t>(k) * nf + f0 + f) * kA + a] = th;
              lt[k * kA + a] = log(clip01(th));
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int a = 0; a < kA; ++a)
              lt[k * kA + a] = log(clip01(t.theta[(static_cast<size_t>(k) * nf + f0 + f) * kA + a]));
        }
      }
      __syncthreads();
      if (e_active) {
        for (int f = ep; f < fn; f += parts) {
          const int xv = xt[f * N + en];
          const double* lt = logt + static_cast<size_t>(f) * K * kA + xv;
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] += lt[k * kA];
        }
      }
      __syncthreads();
    }
    // reduce the parts of every read
    if (e_active) {
#pragma unroll
      for (int k = 0; k < K; ++k) part[static_cast<size_t>(tid) * K + k] = acc[k];
    }
    __syncthreads();
    if (tid < N) {
      double L[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int p = 0; p < parts; ++p) s += part[static_cast<size_t>(p * N + tid) * K + k];
        S[k] = s;
        L[k] = s + log(pis[k]);  // pi is not clamped in the E-step (:150)
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double den = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) den += exp(fmin(fmax(L[k] - L[i], -700.0), 700.0));
        gam[tid * K + i] = 1.0 / den;
      }
    }
    __syncthreads();
  }
  if (tid < N) {
    double lik = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double g = gam[tid * K + k];
      lik += (S[k] + log(clip01(pis[k]))) * g;
      t.gamma[static_cast<size_t>(tid) * K + k] = g;
    }
    t.loglik[tid] = lik;
  }
  if (tid < K) t.pi[tid] = pis[tid];
  if (tid == 0) *t.status = -1;
}

size_t em_smem_bytes(int N, int K, int T, int ft) {
  return (static_cast<size_t>(N) * K + static_cast<size_t>(ft) * K * kA + static_cast<size_t>(T) * K + 2 * K) * sizeof(double) +
         static_cast<size_t>(ft) * N + 16;
}

template <int K, int T>
cudaError_t launch_k(const EmTask* d_tasks, const int32_t* d_ids, int n, size_t smem) {
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(em_kernel<K, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  em_kernel<K, T><<<n, T, smem, cudaStreamPerThread>>>(d_tasks, d_ids);
  return cudaGetLastError();
}

template <int T>
cudaError_t launch(int K, const EmTask* d_tasks, const int32_t* d_ids, int n, size_t smem) {
  switch (K) {
    case 1: return launch_k<1, T>(d_tasks, d_ids, n, smem);
    case 2: return launch_k<2, T>(d_tasks, d_ids, n, smem);
    case 3: return launch_k<3, T>(d_tasks, d_ids, n, smem);
    case 4: return launch_k<4, T>(d_tasks, d_ids, n, smem);
    case 5: return launch_k<5, T>(d_tasks, d_ids, n, smem);
    case 6: return launch_k<6, T>(d_tasks, d_ids, n, smem);
    case 7: return launch_k<7, T>(d_tasks, d_ids, n, smem);
    case 8: return launch_k<8, T>(d_tasks, d_ids, n, smem);
    case 9: return launch_k<9, T>(d_tasks, d_ids, n, smem);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace
}  // namespace svs

using namespace svs;

extern "C" int svs_em_batch(svs_ctx* ctx, int64_t n_tasks, const int8_t* X, const int64_t* x_off,
                            const int32_t* N, const int32_t* nf, const int32_t* K, const int32_t* init_labels,
                            const int64_t* lab_off, int32_t n_steps_default, const int32_t* n_steps,
                            int32_t want_theta, double* gamma, const int64_t* gamma_off, double* theta_io, const int64_t* theta_off,
                            double* pi_io, const int64_t* pi_off, double* loglik, const int64_t* lik_off,
                            int32_t* status) {
  if (!ctx || n_tasks < 0) return fail(ctx, SVS_ERR_ARG, "null argument");
  if (n_tasks == 0) return SVS_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  SVS_CUDA(ctx, cudaSetDevice(ctx->device));
  // sizes: X is shared between the tasks of one window, so x_off may repeat
  size_t x_total = 0, lab_total = 0, g_total = 0, th_total = 0, pi_total = 0, lik_total = 0;
  bool any_from_theta = false;
  for (int64_t t = 0; t < n_tasks; ++t) {
    if (lab_off[t] < 0) any_from_theta = true;
    if (K[t] < 1 || K[t] > 9) return fail(ctx, SVS_ERR_ARG, "K must be 1..9");
    if (N[t] < 1 || N[t] > 1024) return fail(ctx, SVS_ERR_UNSUPPORTED, "mixture model supports 1..1024 reads per window");
    if (nf[t] < 1) return fail(ctx, SVS_ERR_ARG, "nf must be positive");
    x_total = std::max(x_total, static_cast<size_t>(x_off[t]) + static_cast<size_t>(N[t]) * nf[t]);
    if (lab_off[t] >= 0) lab_total = std::max(lab_total, static_cast<size_t>(lab_off[t]) + N[t]);
    g_total = std::max(g_total, static_cast<size_t>(gamma_off[t]) + static_cast<size_t>(N[t]) * K[t]);
    if (want_theta || lab_off[t] < 0)
      th_total = std::max(th_total, static_cast<size_t>(theta_off[t]) + static_cast<size_t>(K[t]) * nf[t] * kA);
    pi_total = std::max(pi_total, static_cast<size_t>(pi_off[t]) + K[t]);
    lik_total = std::max(lik_total, static_cast<size_t>(lik_off[t]) + N[t]);
  }
  int8_t* d_X = nullptr; int32_t *d_lab = nullptr, *d_status = nullptr, *d_ids = nullptr;
  double *d_g = nullptr, *d_th = nullptr, *d_pi = nullptr, *d_lik = nullptr; EmTask* d_tasks = nullptr;
  auto cleanup = [&]() {
    ((d_X) ? cudaFreeAsync(d_X, cudaStreamPerThread) : cudaSuccess); ((d_lab) ? cudaFreeAsync(d_lab, cudaStreamPerThread) : cudaSuccess); ((d_status) ? cudaFreeAsync(d_status, cudaStreamPerThread) : cudaSuccess); ((d_ids) ? cudaFreeAsync(d_ids, cudaStreamPerThread) : cudaSuccess); ((d_g) ? cudaFreeAsync(d_g, cudaStreamPerThread) : cudaSuccess); ((d_th) ? cudaFreeAsync(d_th, cudaStreamPerThread) : cudaSuccess);
    ((d_pi) ? cudaFreeAsync(d_pi, cudaStreamPerThread) : cudaSuccess); ((d_lik) ? cudaFreeAsync(d_lik, cudaStreamPerThread) : cudaSuccess); ((d_tasks) ? cudaFreeAsync(d_tasks, cudaStreamPerThread) : cudaSuccess);
  };
#define SVS_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); \
    return fail(ctx, SVS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_X), x_total + 16, cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_lab), (lab_total + 4) * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_status), n_tasks * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_ids), n_tasks * sizeof(int32_t), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_g), (g_total + 2) * sizeof(double), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_th), (th_total + 2) * sizeof(double), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_pi), (pi_total + 2) * sizeof(double), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_lik), (lik_total + 2) * sizeof(double), cudaStreamPerThread));
  SVS_CU(cudaMallocAsync(reinterpret_cast<void**>(&d_tasks), n_tasks * sizeof(EmTask), cudaStreamPerThread));
  SVS_CU(svs_memcpy_pt(d_X, X, x_total, cudaMemcpyHostToDevice));
  if (lab_total) SVS_CU(svs_memcpy_pt(d_lab, init_labels, lab_total * sizeof(int32_t), cudaMemcpyHostToDevice));
  // start states given as theta/pi travel to the device; the rest is output only
  if ((want_theta || any_from_theta) && !theta_io) { cleanup(); return fail(ctx, SVS_ERR_ARG, "theta buffer required"); }
  for (int64_t t = 0; t < n_tasks && any_from_theta; ++t) {
    if (lab_off[t] >= 0) continue;
    SVS_CU(svs_memcpy_pt(d_th + theta_off[t], theta_io + theta_off[t],
                      static_cast<size_t>(K[t]) * nf[t] * kA * sizeof(double), cudaMemcpyHostToDevice));
  }
  SVS_CU(svs_memcpy_pt(d_pi, pi_io, pi_total * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<EmTask> tasks(n_tasks);
  std::vector<std::vector<int32_t>> by_class(18);  // (K-1)*2 + big
  std::vector<size_t> smem_class(18, 0);
  for (int64_t t = 0; t < n_tasks; ++t) {
    EmTask& e = tasks[t];
    e.X = d_X + x_off[t];
    e.labels = lab_off[t] >= 0 ? d_lab + lab_off[t] : nullptr;
    e.gamma = d_g + gamma_off[t];
    e.theta = (want_theta || lab_off[t] < 0) ? d_th + theta_off[t] : nullptr;
    e.pi = d_pi + pi_off[t];
    e.loglik = d_lik + lik_off[t];
    e.status = d_status + t;
    e.N = N[t]; e.nf = nf[t];
    e.n_steps = n_steps ? n_steps[t] : n_steps_default;
    const int big = N[t] > 256;
    const int T = big ? 1024 : 256;
    int ft = std::min(256, std::max(32, 32768 / N[t] / 32 * 32));
    ft = std::min(ft, (nf[t] + 31) / 32 * 32);
    e.ft = ft;
    const int cls = (K[t] - 1) * 2 + big;
    by_class[cls].push_back(static_cast<int32_t>(t));
    smem_class[cls] = std::max(smem_class[cls], em_smem_bytes(N[t], K[t], T, ft));
  }
  SVS_CU(svs_memcpy_pt(d_tasks, tasks.data(), n_tasks * sizeof(EmTask), cudaMemcpyHostToDevice));
  std::vector<int32_t> ids;
  std::vector<size_t> cls_off(19, 0);
  for (int c = 0; c < 18; ++c) {
    cls_off[c] = ids.size();
    ids.insert(ids.end(), by_class[c].begin(), by_class[c].end());
  }
  cls_off[18] = ids.size();
  SVS_CU(svs_memcpy_pt(d_ids, ids.data(), ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  for (int c = 0; c < 18; ++c) {
    const int n = static_cast<int>(by_class[c].size());
    if (!n) continue;
    if (smem_class[c] > 220 * 1024) { cleanup(); return fail(ctx, SVS_ERR_CAPACITY, "mixture model task exceeds shared memory"); }
    const int Kc = c / 2 + 1;
    if (c & 1) SVS_CU(launch<1024>(Kc, d_tasks, d_ids + cls_off[c], n, smem_class[c]));
    else SVS_CU(launch<256>(Kc, d_tasks, d_ids + cls_off[c], n, smem_class[c]));
  }
  SVS_CU(cudaStreamSynchronize(cudaStreamPerThread));
  SVS_CU(svs_memcpy_pt(gamma, d_g, g_total * sizeof(double), cudaMemcpyDeviceToHost));
  if (want_theta && th_total) SVS_CU(svs_memcpy_pt(theta_io, d_th, th_total * sizeof(double), cudaMemcpyDeviceToHost));
  SVS_CU(svs_memcpy_pt(pi_io, d_pi, pi_total * sizeof(double), cudaMemcpyDeviceToHost));
  SVS_CU(svs_memcpy_pt(loglik, d_lik, lik_total * sizeof(double), cudaMemcpyDeviceToHost));
  SVS_CU(svs_memcpy_pt(status, d_status, n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost));
#undef SVS_CU
  cleanup();
  return SVS_OK;
}
