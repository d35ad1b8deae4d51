// gemb200_params.h — kernel parameter block (passed by value in the kernel-parameter constant bank).
//
// Everything a thread needs besides its own env's state is here, so a launch reads no global-memory parameter
// table: uniform operands come straight from the constant bank (c[0x0][...]) at no issue cost.
#pragma once
#include <stdint.h>

namespace gemb200 {

constexpr int kMaxState = 24;
constexpr int kMaxRef = 4;
constexpr int kMaxConstraints = 4;
constexpr int kMaxX = 6;  // real-typed ODE states per env (omega + motor states without the angle): SCIM 5

// Motor families = template specialisations of the step kernel.
enum MotorFamily : int {
  kDC1 = 0,   // one armature current: PermEx, Series       state [omega, torque, i, u, u_sup]
  kDC2 = 1,   // two currents: Shunt (+i_sum), ExtEx         state [omega, torque, i_a, i_e, u(_a), (u_e,) u_sup(, i_sum)]
  kSYNC = 2,  // PMSM, SynRM                                 14 states
  kEESM = 3,  // 16 states
  kSCIM = 4   // 14 states
};

// RNG stream ids (word 3 of the Philox counter); shared convention with the test oracle.
enum : uint32_t {
  kStreamWalk = 1, kStreamSubep = 2, kStreamInit = 3, kStreamSubepHi = 18,
  kStreamWalkR = 5, kStreamSubepR = 6, kStreamSubepHiR = 22  // "R": draws made right after an in-kernel auto-reset
};

template <typename real>
struct StepParams {
  // ---- batch ----
  int32_t n;              // envs in this handle
  int64_t env_offset;     // global index of env 0 (sharding)
  uint32_t seed_lo, seed_hi;
  uint32_t gstep_lo, gstep_hi;  // unique id of this API call (reset or step): RNG counter words 0,1
  // ---- persistent per-env state (SoA, owned by the handle) ----
  real* x;                // [n_x][n]  omega, currents (, fluxes)
  double* eps;            // [n]       electrical angle, wrapped to (-pi, pi]; nullptr for DC
  real* ref_val;          // [n_ref][n]
  real* ref_sigma;        // [n_ref][n]   (Wiener slots only)
  int32_t* ref_left;      // [n_ref][n]
  uint16_t* sw;           // [n] finite 2QC switching states, 2 bits per leg; nullptr unless finite && interlock
  // ---- I/O of this call (caller-owned) ----
  const void* action;
  real* obs;
  real* ref_out;
  real* reward;
  uint8_t* term;
  const uint8_t* reset_mask;  // reset kernel only
  // ---- system ----
  int32_t motor_kind;     // gemb200_motor_kind (runtime variant inside a family)
  int32_t conv_kind[2];
  int32_t load_kind;
  int32_t solver_kind;
  int32_t nsteps;
  int32_t autoreset;
  int32_t two_segment;    // finite && interlocking_time > 0
  real tau;               // step
  real til;               // interlocking time
  real til_over_tau;
  real u_sup;
  double pole_pairs;      // d eps / dt = p * omega, accumulated in double
  real c[20];             // motor model coefficients (sparse layout per family, see fill_motor_coeffs)
  real tq[4];             // torque coefficients
  real load_a, load_b, load_c, inv_j, omega_lim, omega_lin;
  real inv_lim[kMaxState];
  real init_x[kMaxX];
  double init_eps;
  real reset_obs[kMaxState];  // observation right after a reset (constant initial state)
  // ---- constraint monitor ----
  int32_t n_constraints;
  int32_t con_kind[kMaxConstraints];
  uint32_t con_mask[kMaxConstraints];
  // ---- reward: sum over n_rw terms  w * (|s[idx] - ref| * inv_len)^pow ----
  int32_t n_rw;
  int32_t rw_state[kMaxState];
  int32_t rw_ref[kMaxState];   // reference slot or -1 (reference 0)
  int32_t rw_pow1[kMaxState];  // 1 if power == 1
  real rw_w[kMaxState];
  real rw_inv_len[kMaxState];
  real rw_pow[kMaxState];
  real bias, viol_reward;
  // ---- reference generators ----
  int32_t n_ref;
  int32_t any_wiener;
  int32_t ref_kind[kMaxRef];
  int32_t ref_state[kMaxRef];
  real ref_const[kMaxRef];
  real ref_lo[kMaxRef], ref_hi[kMaxRef];
  real ref_init_lo[kMaxRef], ref_init_span[kMaxRef];
  real ref_lsig_lo[kMaxRef], ref_lsig_span[kMaxRef];  // log10 sigma range
  int32_t ref_len_lo[kMaxRef], ref_len_span[kMaxRef];
};

}  // namespace gemb200
