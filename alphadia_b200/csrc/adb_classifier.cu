// alphadia_b200 — FDR classifier inference on the device (sm_100a).
//
// Replaces BinaryClassifierLegacyNewBatching.predict_proba (alphadia/fdr/classifiers.py:441-470), i.e. FeedForwardNN.forward in
// eval mode (classifiers.py:473-532): BatchNorm1d with its running statistics -> [Linear -> ReLU] per hidden layer (dropout is
// the identity in eval mode) -> Linear -> softmax.  The arithmetic type is float32 like torch's; a contraction of 47 x 100
// and smaller in fp32 has no use for the tensor cores at the 1e-4 parity bar (bf16 / tf32 inputs would break it), so this is a
// plain FFMA kernel: one thread per PSM row, all weights staged in shared memory once per CTA (transposed to [in][out] so
// that the four outputs a thread accumulates at a time are one 128-bit broadcast load), activations ping-pong through a
// per-thread column of shared memory.
#include <cstdio>
#include <string>
#include <vector>

#include "adb_common.cuh"
#include "../../include/alphadia_b200.h"

int adb_set_error(const std::string& msg);  // adb_api.cu

namespace {

constexpr int CL_THREADS = 128;
constexpr int CL_MAX_LAYERS = 8;
constexpr int CL_MAX_WIDTH = 128;  // widest layer (input included)

struct ClNet {
  int input_dim, n_layers;
  int dims[CL_MAX_LAYERS + 1];   // dims[0] = input_dim, dims[l + 1] = output width of layer l
  int w_off[CL_MAX_LAYERS];      // float offsets into the packed parameter block: W^T [in][out_padded], then bias [out_padded]
  int b_off[CL_MAX_LAYERS];
  int out_pad[CL_MAX_LAYERS];    // output width rounded up to 4
  int bn_off;                    // scale[input_dim], shift[input_dim]
  int total;                     // floats in the parameter block
};

__global__ void __launch_bounds__(CL_THREADS) classifier_kernel(ClNet net, const float* __restrict__ params, int64_t n,
                                                                const float* __restrict__ x, float* __restrict__ proba, int width) {
  extern __shared__ float4 smem4[];
  float* sp = (float*)smem4;                       // parameters
  float* act = sp + ((net.total + 3) & ~3);        // [2][width][CL_THREADS]
  for (int t = threadIdx.x; t < net.total; t += CL_THREADS) sp[t] = params[t];
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * CL_THREADS + threadIdx.x;
  if (row >= n) return;
  float* a_in = act + threadIdx.x;
  float* a_out = act + (size_t)width * CL_THREADS + threadIdx.x;
  // BatchNorm1d, eval mode: (x - mean) / sqrt(var + eps) * weight + bias, folded into scale and shift on the host exactly as
  // torch evaluates it is not possible bit for bit, so the four steps are kept: sp[bn + k] = mean, var_eps_rsqrt, weight, bias
  const float* bn = sp + net.bn_off;
  for (int k = 0; k < net.input_dim; k++) {
    const float v = x[row * net.input_dim + k];
    a_in[k * CL_THREADS] = (v - bn[k]) * bn[net.input_dim + k] * bn[2 * net.input_dim + k] + bn[3 * net.input_dim + k];
  }
  for (int l = 0; l < net.n_layers; l++) {
    const int in = net.dims[l], out = net.dims[l + 1], op = net.out_pad[l];
    const float* W = sp + net.w_off[l];
    const float* B = sp + net.b_off[l];
    const bool last = l == net.n_layers - 1;
    for (int j = 0; j < op; j += 4) {
      float4 acc = *(const float4*)(B + j);
      for (int k = 0; k < in; k++) {
        const float a = a_in[k * CL_THREADS];
        const float4 w = *(const float4*)(W + (size_t)k * op + j);
        acc.x = fmaf(w.x, a, acc.x); acc.y = fmaf(w.y, a, acc.y); acc.z = fmaf(w.z, a, acc.z); acc.w = fmaf(w.w, a, acc.w);
      }
      if (!last) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
      a_out[(j + 0) * CL_THREADS] = acc.x;
      if (j + 1 < out) a_out[(j + 1) * CL_THREADS] = acc.y;
      if (j + 2 < out) a_out[(j + 2) * CL_THREADS] = acc.z;
      if (j + 3 < out) a_out[(j + 3) * CL_THREADS] = acc.w;
    }
    float* t = a_in; a_in = a_out; a_out = t;
  }
  // softmax over the output_dim logits (classifiers.py:524-526)
  const int od = net.dims[net.n_layers];
  float mx = a_in[0];
  for (int j = 1; j < od; j++) mx = fmaxf(mx, a_in[j * CL_THREADS]);
  float s = 0.f;
  for (int j = 0; j < od; j++) { const float e = expf(a_in[j * CL_THREADS] - mx); a_in[j * CL_THREADS] = e; s += e; }
  for (int j = 0; j < od; j++) proba[row * od + j] = a_in[j * CL_THREADS] / s;
}

}  // namespace

extern "C" int adb_classifier_predict_proba(int device, const adb_classifier_desc* d, int64_t n, const float* x, float* proba_out) {
  if (!d || (n > 0 && (!x || !proba_out))) return adb_set_error("null argument");
  if (n < 0) return adb_set_error("negative size");
  if (d->n_layers < 1 || d->n_layers > CL_MAX_LAYERS) return adb_set_error("classifier: 1 to 8 linear layers are supported");
  if (d->input_dim < 1 || d->input_dim > CL_MAX_WIDTH) return adb_set_error("classifier: input_dim must be in [1, 128]");
  ClNet net{};
  net.input_dim = d->input_dim;
  net.n_layers = d->n_layers;
  net.dims[0] = d->input_dim;
  int off = 0, width = d->input_dim;
  for (int l = 0; l < d->n_layers; l++) {
    const int out = d->layer_dims[l];
    if (out < 1 || out > CL_MAX_WIDTH) return adb_set_error("classifier: layer widths must be in [1, 128]");
    net.dims[l + 1] = out;
    net.out_pad[l] = (out + 3) & ~3;
    net.w_off[l] = off; off += net.dims[l] * net.out_pad[l];
    net.b_off[l] = off; off += net.out_pad[l];
    width = std::max(width, net.out_pad[l]);
  }
  net.bn_off = off; off += 4 * d->input_dim;
  net.total = off;
  std::vector<float> host((size_t)off, 0.f);
  for (int l = 0; l < d->n_layers; l++) {
    const int in = net.dims[l], out = net.dims[l + 1], op = net.out_pad[l];
    for (int j = 0; j < out; j++) {
      for (int k = 0; k < in; k++) host[(size_t)net.w_off[l] + (size_t)k * op + j] = d->weights[l][(size_t)j * in + k];  // torch: [out, in]
      host[(size_t)net.b_off[l] + j] = d->biases[l][j];
    }
  }
  for (int k = 0; k < d->input_dim; k++) {
    host[(size_t)net.bn_off + k] = d->bn_mean[k];
    host[(size_t)net.bn_off + d->input_dim + k] = 1.0f / sqrtf(d->bn_var[k] + d->bn_eps);
    host[(size_t)net.bn_off + 2 * d->input_dim + k] = d->bn_weight ? d->bn_weight[k] : 1.0f;
    host[(size_t)net.bn_off + 3 * d->input_dim + k] = d->bn_bias ? d->bn_bias[k] : 0.0f;
  }
  const size_t smem = sizeof(float) * (((size_t)off + 3) & ~(size_t)3) + sizeof(float) * 2 * (size_t)width * CL_THREADS;
  if (smem > 220 * 1024) return adb_set_error("classifier: the network does not fit the shared memory of one SM");
  if (n == 0) return 0;
  if (cudaSetDevice(device) != cudaSuccess) return adb_set_error("cudaSetDevice failed");
  const int od = net.dims[net.n_layers];
  float *d_params = nullptr, *d_x = nullptr, *d_p = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_params, sizeof(float) * (size_t)off);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_x, sizeof(float) * (size_t)n * d->input_dim);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_p, sizeof(float) * (size_t)n * od);
  if (e == cudaSuccess) e = cudaMemcpy(d_params, host.data(), sizeof(float) * (size_t)off, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_x, x, sizeof(float) * (size_t)n * d->input_dim, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(classifier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    classifier_kernel<<<(unsigned)((n + CL_THREADS - 1) / CL_THREADS), CL_THREADS, smem>>>(net, d_params, n, d_x, d_p, width);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(proba_out, d_p, sizeof(float) * (size_t)n * od, cudaMemcpyDeviceToHost);
  cudaFree(d_params); cudaFree(d_x); cudaFree(d_p);
  if (e != cudaSuccess) return adb_set_error(std::string("classifier inference failed: ") + cudaGetErrorString(e));
  return 0;
}
