// traffic_probe.cu — calibration: how long does the PMSM step's memory traffic take with NO compute?
// Same streams and access shapes as step_kernel<SYNC,cont,f32,NREF=2,AoS>: per env read 2x float4 (record) + float2 (angle) +
// 3 floats (action, stride 12); write 2x float4 + float2 + 14 floats (row-per-env obs via 16-byte stores) + float2 (ref) +
// float (reward) + uint8 (terminated).  R replicas are rotated so every byte is cold (R*166 MB >> 126 MB L2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o traffic_probe traffic_probe.cu ; run: ./traffic_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

struct Bufs { float4 *st0, *st1; float2* eps; float* act; float4* obs; float2* ref; float* rew; uint8_t* term; };

// V6: the record layout since the hot/cold split — st0 (hot) is written back, st1 + one float (cold, 20 B) only read; PF: L2 prefetch of
// the env `pf` positions ahead, as the step kernel does
template <bool WRITE_OBS, bool WRITE_STATE, bool V6 = false>
__global__ void probe(Bufs b, int n, int pf = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = b.st0[i], c = b.st1[i];
  float2 e = b.eps[i];
  if (V6) a.y += b.rew[i] * 0.f + reinterpret_cast<const float*>(b.ref)[i];  // the 5th cold word (4 B)
  if (pf > 0 && i + pf < n) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b.st0 + i + pf));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b.st1 + i + pf));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b.eps + i + pf));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b.act + (size_t)(i + pf) * 3));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float*>(b.ref) + i + pf));
  }
  const float* ap = b.act + (size_t)i * 3;
  const float s = ap[0] + ap[1] + ap[2];
  a.x += s; c.y += e.x;
  if (WRITE_STATE) { b.st0[i] = a; if (!V6) b.st1[i] = c; b.eps[i] = e; }
  b.ref[i] = make_float2(c.x, c.y);
  b.rew[i] = s;
  b.term[i] = (uint8_t)(s > 100.f);
  if (WRITE_OBS) {
    // each warp owns 32 rows x 14 floats = 112 float4, written as 4 coalesced 16-byte stores per lane (last one half-masked)
    const int lane = threadIdx.x & 31;
    float4* base = b.obs + (size_t)(i - lane) * 14 / 4;
    const float4 v = make_float4(a.x, a.y, c.x, c.y);
#pragma unroll
    for (int it = 0; it < 4; ++it) { const int k = it * 32 + lane; if (k < 112) base[k] = v; }
  }
}

int main() {
  const int n = 1 << 20, R = 4, K = 200;
  std::vector<Bufs> bs(R);
  for (auto& b : bs) {
    cudaMalloc(&b.st0, n * 16); cudaMalloc(&b.st1, n * 16); cudaMalloc(&b.eps, n * 8); cudaMalloc(&b.act, n * 12);
    cudaMalloc(&b.obs, (size_t)n * 56); cudaMalloc(&b.ref, n * 8); cudaMalloc(&b.rew, n * 4); cudaMalloc(&b.term, n);
    cudaMemset(b.st0, 0, n * 16); cudaMemset(b.st1, 0, n * 16); cudaMemset(b.eps, 0, n * 8); cudaMemset(b.act, 0, n * 12);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int mode, int block) {
    const int grid = (n + block - 1) / block;
    for (int w = 0; w < 2; ++w) {
      cudaEventRecord(e0);
      for (int k = 0; k < K; ++k) {
        Bufs& b = bs[k % R];
        if (mode == 0) probe<true, true><<<grid, block>>>(b, n);
        else if (mode == 1) probe<false, true><<<grid, block>>>(b, n);
        else if (mode == 2) probe<true, false><<<grid, block>>>(b, n);
        else if (mode == 3) probe<true, true, true><<<grid, block>>>(b, n, 0);
        else probe<true, true, true><<<grid, block>>>(b, n, 75776);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / K * 1e3f;
  };
  const double rd = 52.0, wr_full = 109.0;
  for (int block : {128, 256}) {
    float t = run(0, block);
    printf("{\"probe\": \"full traffic (52 B read + 109 B written per env)\", \"block\": %d, \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", block, t, (rd + wr_full) * n / t / 1e3);
    t = run(1, block);
    printf("{\"probe\": \"no obs rows (52 B read + 53 B written)\", \"block\": %d, \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", block, t, (rd + 53.0) * n / t / 1e3);
    t = run(2, block);
    printf("{\"probe\": \"no state write-back (52 B read + 69 B written)\", \"block\": %d, \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", block, t, (rd + 69.0) * n / t / 1e3);
    t = run(3, block);
    printf("{\"probe\": \"v6 traffic: hot/cold records (56 B read + 93 B written)\", \"block\": %d, \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", block, t, 149.0 * n / t / 1e3);
    t = run(4, block);
    printf("{\"probe\": \"v6 traffic + L2 prefetch of env i+75776\", \"block\": %d, \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", block, t, 149.0 * n / t / 1e3);
  }
  return 0;
}
