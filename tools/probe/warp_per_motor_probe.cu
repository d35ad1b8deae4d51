// warp_per_motor_probe.cu — experiment behind DESIGN.md §4 "why thread-per-env":
// BASELINE.json's north_star sketches a ONE-WARP-PER-MOTOR layout (state + parameter block staged in shared memory, float4
// loads, warp shuffles for the Clarke/Park transforms).  This probe implements exactly that for the PMSM physics step
// (ContB6 duty -> u_abc -> Clarke (shuffle reduction over the 3 phase lanes) -> Park -> RK4 with one lane per ODE row ->
// 14-entry normalised observation written by 14 lanes) and times it against the same physics with one THREAD per motor.
// Physics only (no reward / reference epilogue) on both sides; N = 2^20 motors, cold buffers rotated (4 replicas).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_per_motor_probe warp_per_motor_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

struct P {  // PMSM model constants (Cont-CC-PMSM-v0 defaults) and limits
  float c0, c1, c2, c3, c4, c5, c6, tq0, tq1, u_sup, tau, inv_lim[14], kang;  // kang = p * tau / (2 pi) [turns per rad/s]
};
struct Bufs { float4* st; float2* eps; const float* act; float* obs; };

__device__ __forceinline__ void rhs(const P& p, float w, float id, float iq, float ud, float uq, float& did, float& diq) {
  did = p.c0 * id + p.c1 * ud + p.c2 * w * iq;
  diq = p.c3 * w + p.c4 * iq + p.c5 * uq + p.c6 * w * id;
}

// ------------------------------------------------------------------ one thread per motor
__global__ void thread_per_motor(Bufs b, P p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 s = b.st[i];  // omega, id, iq, unused
  float2 e = b.eps[i];
  const float* a = b.act + (size_t)i * 3;
  float sn, cs;
  sincospif(2.f * e.x + 2.f * e.y, &sn, &cs);
  float u[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = (fminf(fmaxf(0.5f * (a[k] + 1.f), 0.f), 1.f) - 0.5f) * p.u_sup;
  const float ua = 2.f / 3.f * (u[0] - 0.5f * (u[1] + u[2])), ub = 0.57735027f * (u[1] - u[2]);
  const float ud = cs * ua + sn * ub, uq = -sn * ua + cs * ub;
  float w = s.x, id = s.y, iq = s.z, k1d, k1q, k2d, k2q, k3d, k3q, k4d, k4q;
  const float h = p.tau;
  rhs(p, w, id, iq, ud, uq, k1d, k1q);
  rhs(p, w, id + 0.5f * h * k1d, iq + 0.5f * h * k1q, ud, uq, k2d, k2q);
  rhs(p, w, id + 0.5f * h * k2d, iq + 0.5f * h * k2q, ud, uq, k3d, k3q);
  rhs(p, w, id + h * k3d, iq + h * k3q, ud, uq, k4d, k4q);
  id += h / 6.f * (k1d + 2.f * k2d + 2.f * k3d + k4d);
  iq += h / 6.f * (k1q + 2.f * k2q + 2.f * k3q + k4q);
  const float al = cs * id - sn * iq, be = sn * id + cs * iq;
  float t = e.x + w * p.kang;
  t -= rintf(t);
  float o[14] = {w, (p.tq0 + p.tq1 * id) * iq, al, -0.5f * al + 0.8660254f * be, -0.5f * al - 0.8660254f * be, id, iq, u[0], u[1], u[2], ud, uq, 2.f * t, p.u_sup};
  b.st[i] = make_float4(w, id, iq, s.w);
  b.eps[i] = make_float2(t, e.y);
  float* row = b.obs + (size_t)i * 14;
#pragma unroll
  for (int k = 0; k < 14; ++k) row[k] = o[k] * p.inv_lim[k];  // (the product kernel stages rows in smem for 16-byte stores)
}

// ------------------------------------------------------------------ one warp per motor (north_star sketch)
__global__ void warp_per_motor(Bufs b, P p, int n) {
  __shared__ float sh_state[8][8];  // per warp: omega, id, iq, eps_hi, eps_lo, sin, cos
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + wib; m < n; m += warps_total) {
    // stage the motor's state vector into shared memory (lane 0: float4 state, lane 1: angle)
    if (lane == 0) { const float4 s = b.st[m]; sh_state[wib][0] = s.x; sh_state[wib][1] = s.y; sh_state[wib][2] = s.z; }
    if (lane == 1) { const float2 e = b.eps[m]; sh_state[wib][3] = e.x; sh_state[wib][4] = e.y; float sn, cs; sincospif(2.f * e.x + 2.f * e.y, &sn, &cs); sh_state[wib][5] = sn; sh_state[wib][6] = cs; }
    __syncwarp();
    const float w = sh_state[wib][0], sn = sh_state[wib][5], cs = sh_state[wib][6];
    // lanes 0..2 = phases a, b, c: duty -> phase voltage; Clarke transform = weighted shuffle reduction over the 3 lanes
    float u_ph = 0.f;
    if (lane < 3) u_ph = (fminf(fmaxf(0.5f * (b.act[(size_t)m * 3 + lane] + 1.f), 0.f), 1.f) - 0.5f) * p.u_sup;
    const float wa = lane == 0 ? 2.f / 3.f : (lane < 3 ? -1.f / 3.f : 0.f), wb = lane == 1 ? 0.57735027f : (lane == 2 ? -0.57735027f : 0.f);
    float ua = wa * u_ph, ub = wb * u_ph;
#pragma unroll
    for (int off = 2; off > 0; off >>= 1) { ua += __shfl_down_sync(0xffffffffu, ua, off); ub += __shfl_down_sync(0xffffffffu, ub, off); }
    ua = __shfl_sync(0xffffffffu, ua, 0); ub = __shfl_sync(0xffffffffu, ub, 0);
    const float ud = cs * ua + sn * ub, uq = -sn * ua + cs * ub;
    // lanes 0 and 1 own one ODE row each (d and q current); partner values come through shuffles
    float x = lane == 0 ? sh_state[wib][1] : sh_state[wib][2], acc = 0.f, xt = x;
    const float h = p.tau;
#pragma unroll
    for (int stage = 0; stage < 4; ++stage) {
      const float other = __shfl_xor_sync(0xffffffffu, xt, 1);
      const float k = lane == 0 ? p.c0 * xt + p.c1 * ud + p.c2 * w * other : p.c3 * w + p.c4 * xt + p.c5 * uq + p.c6 * w * other;
      acc += (stage == 0 || stage == 3) ? k : 2.f * k;
      xt = x + (stage == 2 ? h : 0.5f * h) * k;
    }
    x += h / 6.f * acc;
    const float id = __shfl_sync(0xffffffffu, x, 0), iq = __shfl_sync(0xffffffffu, x, 1);
    float t = sh_state[wib][3] + w * p.kang;
    t -= rintf(t);
    const float al = cs * id - sn * iq, be = sn * id + cs * iq;
    const float u0 = __shfl_sync(0xffffffffu, u_ph, 0), u1 = __shfl_sync(0xffffffffu, u_ph, 1), u2 = __shfl_sync(0xffffffffu, u_ph, 2);
    // lanes 0..13 write one observation entry each (one 56-byte contiguous store per warp)
    float o;
    switch (lane) {
      case 0: o = w; break; case 1: o = (p.tq0 + p.tq1 * id) * iq; break; case 2: o = al; break;
      case 3: o = -0.5f * al + 0.8660254f * be; break; case 4: o = -0.5f * al - 0.8660254f * be; break;
      case 5: o = id; break; case 6: o = iq; break; case 7: o = u0; break; case 8: o = u1; break; case 9: o = u2; break;
      case 10: o = ud; break; case 11: o = uq; break; case 12: o = 2.f * t; break; default: o = p.u_sup; break;
    }
    if (lane < 14) b.obs[(size_t)m * 14 + lane] = o * p.inv_lim[lane];
    if (lane == 0) b.st[m] = make_float4(w, id, iq, 0.f);
    if (lane == 1) b.eps[m] = make_float2(t, sh_state[wib][4]);
    __syncwarp();
  }
}

int main() {
  const int n = 1 << 20, R = 4, K = 50;
  P p{};
  const float r_s = 18e-3f, l_d = 0.37e-3f, l_q = 1.2e-3f, psi = 66e-3f, pp = 3.f;
  p.c0 = -r_s / l_d; p.c1 = 1.f / l_d; p.c2 = l_q * pp / l_d; p.c3 = -psi * pp / l_q; p.c4 = -r_s / l_q; p.c5 = 1.f / l_q; p.c6 = -l_d * pp / l_q;
  p.tq0 = 1.5f * pp * psi; p.tq1 = 1.5f * pp * (l_d - l_q); p.u_sup = 300.f; p.tau = 1e-4f; p.kang = pp * 1e-4f / 6.2831853f;
  const float lim[14] = {418.879f, 160.61f, 400, 400, 400, 400, 400, 150, 150, 150, 150, 150, 1.f, 300};
  for (int k = 0; k < 14; ++k) p.inv_lim[k] = 1.f / lim[k];
  std::vector<Bufs> bs(R);
  for (auto& b : bs) {
    float* act;
    cudaMalloc(&b.st, n * 16); cudaMalloc(&b.eps, n * 8); cudaMalloc(&act, n * 12); cudaMalloc(&b.obs, (size_t)n * 56);
    cudaMemset(b.st, 0, n * 16); cudaMemset(b.eps, 0, n * 8); cudaMemset(act, 0, n * 12);
    b.act = act;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time_it = [&](int which) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      for (int k = 0; k < K; ++k) {
        if (which == 0) thread_per_motor<<<(n + 127) / 128, 128>>>(bs[k % R], p, n);
        else warp_per_motor<<<148 * 8, 256>>>(bs[k % R], p, n);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms / K * 1e3f;
  };
  const float t_thread = time_it(0), t_warp = time_it(1);
  printf("{\"probe\": \"PMSM physics step, N=2^20\", \"thread_per_motor_us\": %.1f, \"warp_per_motor_us\": %.1f, \"ratio\": %.1f, "
         "\"thread_per_motor_steps_per_s\": %.3e, \"warp_per_motor_steps_per_s\": %.3e}\n",
         t_thread, t_warp, t_warp / t_thread, n / (t_thread * 1e-6), n / (t_warp * 1e-6));
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
  return 0;
}
