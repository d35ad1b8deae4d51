// gemb200.cu — C-ABI implementation (include/gemb200.h): handle management, host-side derivation of the model
// constants from the physical parameters (the reference's *_update_model methods), kernel dispatch.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <limits>
#include <string>
#include <vector>

#include "gemb200_launch.cuh"

using namespace gemb200;

// ----------------------------------------------------------------------------------------------------------------
// error reporting
// ----------------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return fail(GEMB200_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ----------------------------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------------------------
struct gemb200_handle {
  gemb200_config cfg;
  int fam = 0, n_state = 0, n_ode = 0, n_act = 0, n_ref = 0, nx = 0;
  bool has_eps = false, any_wiener = false, two_segment = false, any_switched = false;
  size_t rsz = 4;  // sizeof(real)
  // persistent device state
  int NH = 0, NC = 0;  // words per env in the hot / cold record
  void* d_st = nullptr;
  void* d_stc = nullptr;
  double* d_eps = nullptr;
  uint16_t* d_sw = nullptr;
  void* d_fifo = nullptr;
  int fifo_dim = 0;
  void* d_sup = nullptr;   // RC supply state [2][n]
  double* d_supph = nullptr;  // AC supply phase [n]
  void* d_ext = nullptr;       // external speed profile table (real)
  uint32_t* d_kenv = nullptr;  // steps since the reset per env (external speed profile)
  uint32_t* d_swst = nullptr;  // switched reference generators [n_ref][2][n]
  void* d_obsv = nullptr;  // FluxObserver integrator [4][n]: re, im, compensation of re, of im
  void* d_envp = nullptr;    // per-env model coefficients [kCoefWords][n] (gemb200_set_env_params), nullptr: shared coefficients
  int plain_shape = 0;       // the configuration has the PLAIN shape (before per-env parameters switch the specialisation off)
  void* d_imprev = nullptr;  // induction motors with random initial states [2][n]: initial currents of the env's previous episode
  int n_obs = 0, row_stride = 0;
  StepParams<float> pf;
  StepParams<double> pd;
  uint64_t gstep = 0;
  uint64_t n_steps = 0;  // step calls so far (dead-time ring position)
  uint64_t ext_hash = 0; // FNV-1a of the external speed profile table (part of the checkpoint fingerprint)
  uint32_t* d_clock = nullptr;  // device-resident clock (gemb200_set_device_clock): {call id lo, hi, step count lo, ring position, step count hi, -, -, -}
  bool dev_clock = false;       // launches read d_clock instead of gstep / n_steps (which are then stale until the clock is pulled back)
  int64_t launches = 0;
  // host-buffer path
  cudaStream_t hstream = nullptr;
  cudaStream_t hpipe[3] = {nullptr, nullptr, nullptr};
  void *d_act = nullptr, *d_obs = nullptr, *d_ref = nullptr, *d_rew = nullptr;
  uint8_t *d_term = nullptr, *d_mask = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

// ----------------------------------------------------------------------------------------------------------------
// dimensions / validation (SCMLSystem._set_indices physical_systems.py:141-162, :462-485, :594-617, :737-763)
// ----------------------------------------------------------------------------------------------------------------
static bool is_qc(int k) { return k == GEMB200_CONV_1QC || k == GEMB200_CONV_2QC || k == GEMB200_CONV_4QC; }

struct Dims { int fam, n_state, n_ode, n_act, nx; bool has_eps; int n_obs = 0; bool has_observer = false; };  // n_state: the system's own vector, n_obs: after the wrappers

static int derive_dims(const gemb200_config* c, Dims* d) {
  const int k0 = c->converter_kind[0], k1 = c->converter_kind[1];
  switch (c->motor_kind) {
    case GEMB200_MOTOR_PERMEX_DC:
    case GEMB200_MOTOR_SERIES_DC: *d = {kDC1, 5, 2, 1, 2, false}; break;
    case GEMB200_MOTOR_SHUNT_DC: *d = {kDC2, 7, 3, 1, 3, false}; break;
    case GEMB200_MOTOR_EXTEX_DC: *d = {kDC2, 7, 3, 2, 3, false}; break;
    case GEMB200_MOTOR_PMSM:
    case GEMB200_MOTOR_SYNRM: *d = {kSYNC, 14, 4, c->finite ? 1 : 3, 3, true}; break;
    case GEMB200_MOTOR_EESM: *d = {kEESM, 16, 5, c->finite ? 2 : 4, 4, true}; break;
    case GEMB200_MOTOR_SCIM: *d = {kSCIM, 14, 6, c->finite ? 1 : 3, 5, true}; break;
    case GEMB200_MOTOR_DFIM: *d = {kDFIM, 24, 6, c->finite ? 2 : 6, 5, true}; break;
    default: return fail(GEMB200_E_INVALID, "unknown motor_kind");
  }
  const bool three_phase = d->fam >= kSYNC;
  if (c->action_dq) {
    if (!three_phase || c->finite) return fail(GEMB200_E_INVALID, "dq actions need a three-phase motor with a continuous converter");
    d->n_act = c->motor_kind == GEMB200_MOTOR_EESM ? 3 : (c->motor_kind == GEMB200_MOTOR_DFIM ? 4 : 2);
  }
  // state-vector wrappers (gemb200_state_op): width bookkeeping as in the wrappers' set_physical_system
  d->n_obs = d->n_state;
  if (c->n_state_ops < 0 || c->n_state_ops > GEMB200_MAX_STATE_OPS) return fail(GEMB200_E_INVALID, "n_state_ops out of range");
  for (int k = 0; k < c->n_state_ops; ++k) {
    switch (c->sop_kind[k]) {
      case GEMB200_SOP_COS_SIN:
        if (c->sop_idx[k][0] < 0 || c->sop_idx[k][0] >= d->n_obs) return fail(GEMB200_E_INVALID, "CosSinProcessor: angle index out of range");
        d->n_obs += c->sop_idx[k][1] ? 1 : 2;
        break;
      case GEMB200_SOP_FLUX_OBSERVER:
        if (c->motor_kind != GEMB200_MOTOR_SCIM && c->motor_kind != GEMB200_MOTOR_DFIM) return fail(GEMB200_E_INVALID, "FluxObserver needs an induction motor (flux_observer.py:57-60)");
        if (d->has_observer) return fail(GEMB200_E_INVALID, "only one FluxObserver per system");
        for (int q = 0; q < 4; ++q)
          if (c->sop_idx[k][q] < 0 || c->sop_idx[k][q] >= d->n_obs) return fail(GEMB200_E_INVALID, "FluxObserver: state index out of range");
        if (!(c->sop_param[k][3] > 0)) return fail(GEMB200_E_INVALID, "FluxObserver: psi_limit must be positive");
        d->has_observer = true;
        d->n_obs += 2;
        break;
      case GEMB200_SOP_CURRENT_SUM:
        if (c->sop_mask[k] == 0 || (d->n_obs < 32 && (c->sop_mask[k] >> d->n_obs) != 0)) return fail(GEMB200_E_INVALID, "CurrentSumProcessor: state indices out of range");
        d->n_obs += 1;
        break;
      case GEMB200_SOP_NOISE:
        if (c->sop_idx[k][0] < GEMB200_NOISE_NORMAL || c->sop_idx[k][0] > GEMB200_NOISE_LAPLACE) return fail(GEMB200_E_INVALID, "StateNoiseProcessor: unknown distribution");
        if (d->n_obs < 32 && (c->sop_mask[k] >> d->n_obs) != 0) return fail(GEMB200_E_INVALID, "StateNoiseProcessor: state index out of range");
        break;
      default: return fail(GEMB200_E_INVALID, "unknown state op");
    }
    if (d->n_obs > GEMB200_MAX_STATE) return fail(GEMB200_E_INVALID, "state vector too long");
  }
  if (c->action_dq == 2 && !(c->motor_kind == GEMB200_MOTOR_SCIM && d->has_observer))
    return fail(GEMB200_E_INVALID, "action_dq = 2 (observer angle) needs a SCIM with a FluxObserver");
  if (c->motor_kind == GEMB200_MOTOR_DFIM && c->action_dq && !(c->action_dq == 3 && d->has_observer))
    return fail(GEMB200_E_INVALID, "DFIM dq actions are action_dq = 3 and need a FluxObserver (dq_to_abc_action_processor.py:108-137)");
  if (c->action_dq == 3 && c->motor_kind != GEMB200_MOTOR_DFIM) return fail(GEMB200_E_INVALID, "action_dq = 3 is the DFIM processor");
  if (three_phase) {
    if (k0 != GEMB200_CONV_B6) return fail(GEMB200_E_INVALID, "three-phase motors need a B6 bridge in converter slot 0");
    if (c->motor_kind == GEMB200_MOTOR_EESM) {
      if (!is_qc(k1)) return fail(GEMB200_E_INVALID, "EESM needs a 1QC/2QC/4QC excitation converter in slot 1");
    } else if (c->motor_kind == GEMB200_MOTOR_DFIM) {
      if (k1 != GEMB200_CONV_B6) return fail(GEMB200_E_INVALID, "DFIM needs a second B6 bridge (rotor) in converter slot 1");
    } else if (k1 != GEMB200_CONV_NONE) return fail(GEMB200_E_INVALID, "converter slot 1 must be NONE for this motor");
  } else {
    if (!is_qc(k0)) return fail(GEMB200_E_INVALID, "DC motors need a 1QC/2QC/4QC converter in slot 0");
    if (c->motor_kind == GEMB200_MOTOR_EXTEX_DC) {
      if (!is_qc(k1)) return fail(GEMB200_E_INVALID, "ExtEx DC motor needs an excitation converter in slot 1");
    } else if (k1 != GEMB200_CONV_NONE) return fail(GEMB200_E_INVALID, "converter slot 1 must be NONE for this motor");
  }
  return GEMB200_OK;
}

static int validate(const gemb200_config* c) {
  if (!c) return fail(GEMB200_E_INVALID, "config is NULL");
  if (c->struct_size != (int32_t)sizeof(gemb200_config) || c->abi_version != GEMB200_ABI_VERSION)
    return fail(GEMB200_E_ABI, "gemb200_config struct_size/abi_version mismatch (use gemb200_config_init)");
  if (c->n_envs < 1) return fail(GEMB200_E_INVALID, "n_envs must be >= 1");
  if (c->dtype != GEMB200_F32 && c->dtype != GEMB200_F64) return fail(GEMB200_E_INVALID, "bad dtype");
  if (c->layout != GEMB200_LAYOUT_AOS && c->layout != GEMB200_LAYOUT_SOA) return fail(GEMB200_E_INVALID, "bad layout");
  if (c->solver_kind != GEMB200_SOLVER_EULER && c->solver_kind != GEMB200_SOLVER_RK4)
    return fail(GEMB200_E_INVALID, "solver_kind must be EULER or RK4 (the scipy solvers of the reference map to RK4 sub-stepping, see DESIGN.md)");
  if (c->solver_nsteps < 1 || c->solver_nsteps > 1024) return fail(GEMB200_E_INVALID, "solver_nsteps out of range");
  if (!(c->tau > 0)) return fail(GEMB200_E_INVALID, "tau must be positive");
  if (c->interlocking_time < 0 || c->interlocking_time >= c->tau) return fail(GEMB200_E_INVALID, "interlocking_time must be in [0, tau)");
  if (c->interlocking_time1 >= c->tau) return fail(GEMB200_E_INVALID, "interlocking_time1 must be < tau (negative: same as interlocking_time)");
  if (c->load_kind < GEMB200_LOAD_CONST_SPEED || c->load_kind > GEMB200_LOAD_EXT_SPEED) return fail(GEMB200_E_INVALID, "bad load_kind");
  if (c->load_kind == GEMB200_LOAD_EXT_SPEED) {
    if (!c->ext_speed_table || c->ext_speed_len < 4 * c->solver_nsteps + 2) return fail(GEMB200_E_INVALID, "external speed load: table missing or shorter than two steps");
    if (!(c->load_param[GEMB200_LP_TAU_LOAD] > 0)) return fail(GEMB200_E_INVALID, "external speed load: tau_load must be positive");
    if (c->finite && (c->interlocking_time > 0 || c->interlocking_time1 > 0)) return fail(GEMB200_E_INVALID, "external speed load with two-segment steps (finite converter + interlocking time) is not supported: the segment times are off the table grid");
  }
  if (c->n_ref < 0 || c->n_ref > GEMB200_MAX_REF) return fail(GEMB200_E_INVALID, "n_ref out of range");
  if (c->dead_time_steps < 0 || c->dead_time_steps > GEMB200_MAX_DEAD_TIME) return fail(GEMB200_E_INVALID, "dead_time_steps out of range");
  const bool induction = c->motor_kind == GEMB200_MOTOR_SCIM || c->motor_kind == GEMB200_MOTOR_DFIM;
  if (c->init_random && induction && !c->init_im_valid)
    return fail(GEMB200_E_INVALID, "random initial states of an induction motor need init_im (flux-limit constants, see gemb200.h)");
  if (c->init_im_valid && !(induction && c->init_random)) return fail(GEMB200_E_INVALID, "init_im is for induction motors with init_random");
  if (c->init_im_valid && !(c->init_im[4] != 0.0)) return fail(GEMB200_E_INVALID, "init_im[4] (p * l_m / l_r) must be non-zero");
  for (int j = 0; j < GEMB200_MAX_ODE; ++j)
    if (c->init_random && c->init_dist[j] && !(c->init_sigma[j] > 0 && (c->init_hi[j] > c->init_lo[j])))
      return fail(GEMB200_E_INVALID, "truncated-normal initial state needs sigma > 0 and a non-empty interval");
  if (c->supply_kind < GEMB200_SUPPLY_IDEAL || c->supply_kind > GEMB200_SUPPLY_AC1) return fail(GEMB200_E_INVALID, "bad supply_kind");
  if (c->supply_kind == GEMB200_SUPPLY_AC1 && !(c->supply_param[0] > 0)) return fail(GEMB200_E_INVALID, "AC supply needs a positive frequency");
  if (c->supply_kind == GEMB200_SUPPLY_RC && !(c->supply_param[0] > 0 && c->supply_param[1] > 0)) return fail(GEMB200_E_INVALID, "RC supply needs R > 0 and C > 0");
  if (c->n_constraints < 0 || c->n_constraints > GEMB200_MAX_CONSTRAINTS) return fail(GEMB200_E_INVALID, "n_constraints out of range");
  for (int i = 0; i < c->n_constraints; ++i)
    if (c->constraint_kind[i] != GEMB200_CONSTRAINT_LIMIT && c->constraint_kind[i] != GEMB200_CONSTRAINT_SQUARED) return fail(GEMB200_E_INVALID, "bad constraint_kind");
  if (c->autoreset != GEMB200_AUTORESET_NONE && c->autoreset != GEMB200_AUTORESET_SAME_STEP) return fail(GEMB200_E_INVALID, "bad autoreset mode");
  Dims d;
  int rc = derive_dims(c, &d);
  if (rc) return rc;
  if (c->finite && (c->interlocking_time > 0 || c->interlocking_time1 > 0) && c->motor_kind == GEMB200_MOTOR_EESM)
    return fail(GEMB200_E_INVALID, "finite EESM with interlocking time: the reference raises in this configuration "
                                   "(physical_systems.py:632 slices u_in[:2]); not supported");
  int n_entries = c->n_ref;  // parameter entries in use: the output slots plus the extra sub-generators of switched slots
  for (int r = 0; r < c->n_ref; ++r) {
    if (c->ref_sw_count[r] <= 1) continue;
    const int first = c->ref_sw_first[r], cnt = c->ref_sw_count[r];
    if (first < 0 || first + cnt > GEMB200_MAX_REF_ENTRIES) return fail(GEMB200_E_INVALID, "switched reference generator: parameter entries out of range");
    if (first + cnt > n_entries) n_entries = first + cnt;
    if (c->ref_sw_len_lo[r] < 1 || c->ref_sw_len_hi[r] <= c->ref_sw_len_lo[r]) return fail(GEMB200_E_INVALID, "switched reference generator: bad super-episode length range");
    for (int m = 0; m < cnt; ++m) {
      const int k = c->ref_kind[first + m];
      if (k == GEMB200_REF_EXTERNAL) return fail(GEMB200_E_INVALID, "switched reference generator: external sub-generators are not supported");
      if (!(c->ref_sw_cdf[first + m] > 0 && c->ref_sw_cdf[first + m] <= 1.0 + 1e-12)) return fail(GEMB200_E_INVALID, "switched reference generator: bad probabilities");
    }
  }
  for (int r = 0; r < c->n_ref; ++r)
    if (c->ref_state[r] < 0 || c->ref_state[r] >= d.n_obs) return fail(GEMB200_E_INVALID, "ref_state out of range");
  for (int r = 0; r < n_entries; ++r) {
    if (c->ref_kind[r] < GEMB200_REF_CONST || c->ref_kind[r] > GEMB200_REF_TRIANGULAR) return fail(GEMB200_E_INVALID, "bad ref_kind");
    const bool subep = c->ref_kind[r] == GEMB200_REF_WIENER || c->ref_kind[r] >= GEMB200_REF_LAPLACE;
    const bool walk = c->ref_kind[r] == GEMB200_REF_WIENER || c->ref_kind[r] == GEMB200_REF_LAPLACE;
    if (subep && (c->ref_len_lo[r] < 1 || c->ref_len_hi[r] < c->ref_len_lo[r] || c->ref_len_hi[r] > (1 << 24)))
      return fail(GEMB200_E_INVALID, "bad sub-episode length range");
    if (walk && !(c->ref_sigma_lo[r] > 0 && c->ref_sigma_hi[r] >= c->ref_sigma_lo[r])) return fail(GEMB200_E_INVALID, "bad sigma range");
    if (c->ref_kind[r] >= GEMB200_REF_SINUS && !(c->ref_freq_lo[r] > 0 && c->ref_freq_hi[r] >= c->ref_freq_lo[r] && c->ref_amp_lo[r] >= 0))
      return fail(GEMB200_E_INVALID, "bad amplitude / frequency range of a periodic reference generator");
  }
  for (int j = 0; j < d.n_obs; ++j)
    if (c->reward_weight[j] != 0.0 && !(c->state_length[j] > 0)) return fail(GEMB200_E_INVALID, "state_length must be positive for every weighted state");
  for (int j = 0; j < d.n_state; ++j)
    if (!(c->limits[j] != 0.0) && !(c->motor_kind == GEMB200_MOTOR_SHUNT_DC && j == 6)) return fail(GEMB200_E_INVALID, "limits must be non-zero");
  const double j_total = c->load_param[GEMB200_LP_J_LOAD] + c->motor_param[GEMB200_MP_J_ROTOR];
  if (c->load_kind == GEMB200_LOAD_POLY_STATIC && !(j_total > 0)) return fail(GEMB200_E_INVALID, "total inertia must be positive");
  if ((int64_t)c->n_envs * (int64_t)(hot_words(d.nx, c->n_ref) + cold_words(d.nx, c->n_ref) > d.n_obs ? hot_words(d.nx, c->n_ref) + cold_words(d.nx, c->n_ref) : d.n_obs) >= (int64_t)1 << 31)
    return fail(GEMB200_E_INVALID, "n_envs too large for 32-bit element indexing in one handle; shard the batch");
  return GEMB200_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// model constants (double) -> StepParams<real>
// ----------------------------------------------------------------------------------------------------------------
struct Derived {
  double c[20] = {0};
  double tq[4] = {0};
  double reset_obs[GEMB200_MAX_STATE] = {0};
  double reset_obs_du[GEMB200_MAX_STATE] = {0};  // d reset_obs / d u_sup (the voltage entries are linear in u_sup)
  double inv_j = 0, omega_lim = 0, omega_lin = 0;
};

static void derive_model(const gemb200_config* cfg, const Dims& dm, Derived* o) {
  const double* mp = cfg->motor_param;
  const double p = mp[GEMB200_MP_P], r_s = mp[GEMB200_MP_R_S], l_d = mp[GEMB200_MP_L_D], l_q = mp[GEMB200_MP_L_Q];
  double* c = o->c;
  switch (cfg->motor_kind) {
    case GEMB200_MOTOR_PERMEX_DC: {  // dc_permanently_excited_motor.py:71-75
      const double l_a = mp[GEMB200_MP_L_A];
      c[0] = -mp[GEMB200_MP_PSI_E] / l_a; c[1] = -mp[GEMB200_MP_R_A] / l_a; c[2] = 0; c[3] = 1.0 / l_a;
      o->tq[0] = mp[GEMB200_MP_PSI_E]; o->tq[1] = 0;
    } break;
    case GEMB200_MOTOR_SERIES_DC: {  // dc_series_motor.py:66-74
      const double l = mp[GEMB200_MP_L_A] + mp[GEMB200_MP_L_E];
      c[0] = 0; c[1] = (-mp[GEMB200_MP_R_A] - mp[GEMB200_MP_R_E]) / l; c[2] = -mp[GEMB200_MP_L_E_PRIME] / l; c[3] = 1.0 / l;
      o->tq[0] = 0; o->tq[1] = mp[GEMB200_MP_L_E_PRIME];
    } break;
    case GEMB200_MOTOR_SHUNT_DC:
    case GEMB200_MOTOR_EXTEX_DC: {  // dc_motor.py:95-108
      const double l_a = mp[GEMB200_MP_L_A], l_e = mp[GEMB200_MP_L_E];
      c[0] = -mp[GEMB200_MP_R_A] / l_a; c[1] = -mp[GEMB200_MP_L_E_PRIME] / l_a; c[2] = 1.0 / l_a;
      c[3] = -mp[GEMB200_MP_R_E] / l_e; c[4] = 1.0 / l_e;
      o->tq[0] = mp[GEMB200_MP_L_E_PRIME];
    } break;
    case GEMB200_MOTOR_PMSM:
    case GEMB200_MOTOR_SYNRM: {  // permanent_magnet_synchronous_motor.py:107-139, synchronous_reluctance_motor.py:117-139
      const double psi_p = cfg->motor_kind == GEMB200_MOTOR_PMSM ? mp[GEMB200_MP_PSI_P] : 0.0;
      c[0] = -r_s / l_d; c[1] = 1.0 / l_d; c[2] = l_q * p / l_d;
      c[3] = -psi_p * p / l_q; c[4] = -r_s / l_q; c[5] = 1.0 / l_q; c[6] = -l_d * p / l_q;
      o->tq[0] = 1.5 * p * psi_p; o->tq[1] = 1.5 * p * (l_d - l_q);
    } break;
    case GEMB200_MOTOR_EESM: {  // externally_excited_synchronous_motor.py:125-153, :200-203
      const double k = mp[GEMB200_MP_K], r_e = mp[GEMB200_MP_R_E], l_m = mp[GEMB200_MP_L_M], l_e = mp[GEMB200_MP_L_E];
      const double r_E = k * k * 1.5 * r_e, l_M = k * 1.5 * l_m, l_E = k * k * 1.5 * l_e, ik = 2.0 / 3.0 / k;
      const double sigma = 1.0 - l_M * l_M / (l_d * l_E);
      c[0] = (-r_s / sigma) / l_d; c[1] = (l_M * r_E / (sigma * l_E) * ik) / l_d; c[2] = (1.0 / sigma) / l_d;
      c[3] = (-l_M * k / (sigma * l_E)) / l_d; c[4] = (l_q * p / sigma) / l_d;
      c[5] = -r_s / l_q; c[6] = 1.0 / l_q; c[7] = -l_d * p / l_q; c[8] = -p * l_M * ik / l_q;
      const double s2 = l_E * ik;
      c[9] = (l_M * r_s / (sigma * l_d)) / s2; c[10] = (-r_E / sigma * ik) / s2; c[11] = (-l_M / (sigma * l_d)) / s2;
      c[12] = (k / sigma) / s2; c[13] = (-p * l_M * l_q / (sigma * l_d)) / s2;
      o->tq[0] = 1.5 * p * l_M * ik; o->tq[1] = 1.5 * p * (l_d - l_q);
    } break;
    case GEMB200_MOTOR_DFIM:
    case GEMB200_MOTOR_SCIM: {  // induction_motor.py:287-310, :236-249
      const double l_m = mp[GEMB200_MP_L_M], r_r = mp[GEMB200_MP_R_E];
      const double l_s = l_m + mp[GEMB200_MP_L_SIGS], l_r = l_m + mp[GEMB200_MP_L_SIGR];
      const double sigma = (l_s * l_r - l_m * l_m) / (l_s * l_r);
      const double tau_r = l_r / r_r, tau_sig = sigma * l_s / (r_s + r_r * (l_m * l_m) / (l_r * l_r));
      c[0] = -1.0 / tau_sig; c[1] = l_m * r_r / (sigma * l_s * l_r * l_r); c[2] = l_m * p / (sigma * l_r * l_s);
      c[3] = 1.0 / (sigma * l_s); c[4] = l_m / tau_r; c[5] = -1.0 / tau_r; c[6] = p;
      c[7] = -l_m / (sigma * l_r * l_s);  // rotor-voltage column of the current rows (DFIM)
      c[8] = 1.0 / l_r; c[9] = l_m / l_r;  // rotor current i_r = psi_r / l_r - l_m / l_r * i_s (physical_systems.py:946-956)
      o->tq[0] = 1.5 * p * l_m / l_r;
    } break;
  }
  // MechanicalLoad.set_j_rotor mechanical_load.py:188-193, polynomial_static_load.py:60-64
  const double* lp = cfg->load_param;
  const double j_total = lp[GEMB200_LP_J_LOAD] + mp[GEMB200_MP_J_ROTOR];
  o->inv_j = j_total > 0 ? 1.0 / j_total : 0.0;
  o->omega_lin = j_total / lp[GEMB200_LP_TAU_DECAY];
  o->omega_lim = j_total > 0 ? lp[GEMB200_LP_A] / j_total * lp[GEMB200_LP_TAU_DECAY] : 0.0;

  // observation right after reset (SCMLSystem.reset physical_systems.py:256-287, :527-561, :659-693, :816-847) for the
  // constant initial state; converter.reset() gives 0 per QC and -0.5 per B6 leg (converters.py:45-54, :880-886)
  const double* y = cfg->init_ode;
  const double U = cfg->u_sup;
  double* s = o->reset_obs;
  int n = 0;
  s[n++] = y[0];
  switch (dm.fam) {
    case kDC1: s[n++] = (o->tq[0] + o->tq[1] * y[1]) * y[1]; s[n++] = y[1]; s[n++] = 0.0; break;
    case kDC2:
      s[n++] = o->tq[0] * y[1] * y[2]; s[n++] = y[1]; s[n++] = y[2]; s[n++] = 0.0;
      if (cfg->motor_kind == GEMB200_MOTOR_EXTEX_DC) s[n++] = 0.0;
      break;
    default: {
      double eps = y[dm.nx];
      if (eps > M_PI) eps -= 2 * M_PI;
      const double ua = -0.5 * U;
      // abc -> alpha/beta of (ua,ua,ua) is mathematically 0 (the reference shows ~1e-17 round-off here)
      const double ualpha = 2.0 / 3.0 * (ua - 0.5 * ua - 0.5 * ua), ubeta = 2.0 / 3.0 * (0.5 * std::sqrt(3.0) * ua - 0.5 * std::sqrt(3.0) * ua);
      double cs, sn, tqv, ia, ib;
      if (dm.fam == kSCIM || dm.fam == kDFIM) {
        const double ef = std::atan2(y[4], y[3]);
        cs = std::cos(ef); sn = std::sin(ef);
        tqv = o->tq[0] * (y[3] * y[2] - y[4] * y[1]);
        ia = y[1]; ib = y[2];
      } else {
        cs = std::cos(eps); sn = std::sin(eps);
        tqv = dm.fam == kSYNC ? (o->tq[0] + o->tq[1] * y[1]) * y[2] : (o->tq[0] * y[3] + o->tq[1] * y[1]) * y[2];
        ia = cs * y[1] - sn * y[2]; ib = sn * y[1] + cs * y[2];
      }
      const double ud = cs * ualpha + sn * ubeta, uq = -sn * ualpha + cs * ubeta;
      s[n++] = tqv;
      s[n++] = ia; s[n++] = -0.5 * ia + 0.5 * std::sqrt(3.0) * ib; s[n++] = -0.5 * ia - 0.5 * std::sqrt(3.0) * ib;
      if (dm.fam == kDFIM) {
        // physical_systems.py:1062-1113: i_sdq in the field frame; i_rdq (sic) with the angle eps_field - eps_el, i_rdef = its inverse;
        // all six bridge legs at -0.5 u_sup, whose alpha-beta image is 0
        const double ira = o->c[8] * y[3] - o->c[9] * y[1], irb = o->c[8] * y[4] - o->c[9] * y[2];
        const double cfe = std::cos(std::atan2(y[4], y[3]) - eps), sfe = std::sin(std::atan2(y[4], y[3]) - eps);
        s[n++] = cs * y[1] + sn * y[2]; s[n++] = -sn * y[1] + cs * y[2];
        s[n++] = ira; s[n++] = -0.5 * ira + 0.5 * std::sqrt(3.0) * irb; s[n++] = -0.5 * ira - 0.5 * std::sqrt(3.0) * irb;
        s[n++] = cfe * ira + sfe * irb; s[n++] = -sfe * ira + cfe * irb;
        s[n++] = ua; s[n++] = ua; s[n++] = ua; s[n++] = ud; s[n++] = uq;
        s[n++] = ua; s[n++] = ua; s[n++] = ua; s[n++] = cfe * ualpha + sfe * ubeta; s[n++] = -sfe * ualpha + cfe * ubeta;
        s[n++] = eps;
        break;
      }
      if (dm.fam == kSCIM) { s[n++] = cs * y[1] + sn * y[2]; s[n++] = -sn * y[1] + cs * y[2]; }
      else { s[n++] = y[1]; s[n++] = y[2]; }
      if (dm.fam == kEESM) {
        // reference quirk (:659-693): u_abc has 4 entries [ua,ub,uc,u_e=0], then u_dq -> the slots named
        // u_sd,u_sq,u_e receive (0, u_d, u_q)
        s[n++] = y[3];
        s[n++] = ua; s[n++] = ua; s[n++] = ua; s[n++] = 0.0; s[n++] = ud; s[n++] = uq;
      } else {
        s[n++] = ua; s[n++] = ua; s[n++] = ua; s[n++] = ud; s[n++] = uq;
      }
      s[n++] = eps;
    } break;
  }
  s[n++] = U;
  for (int j = 0; j < n; ++j) s[j] /= cfg->limits[j];
  if (cfg->motor_kind == GEMB200_MOTOR_SHUNT_DC) s[n] = s[2] + s[3];
}

template <typename real>
static void fill_params(const gemb200_handle* h, const Dims& dm, const Derived& dv, StepParams<real>* p) {
  const gemb200_config& c = h->cfg;
  std::memset(p, 0, sizeof(*p));
  p->n = c.n_envs;
  p->env_begin = 0; p->env_end = c.n_envs;
  p->env_offset = c.env_index_offset;
  p->seed_lo = (uint32_t)c.seed; p->seed_hi = (uint32_t)(c.seed >> 32);
  for (int r = 0; r < 10; ++r) { p->rk[r][0] = p->seed_lo + (uint32_t)r * 0x9E3779B9u; p->rk[r][1] = p->seed_hi + (uint32_t)r * 0xBB67AE85u; }
  p->st = static_cast<real*>(h->d_st);
  p->stc = static_cast<real*>(h->d_stc);
  p->kstep = 0;
  p->eps = h->d_eps;
  p->sw = h->d_sw;
  p->layout = c.layout;
  p->n_act = dm.n_act;
  p->fifo = static_cast<real*>(h->d_fifo);
  p->action_dq = c.action_dq;
  p->dead_steps = c.dead_time_steps; p->dead_outer = c.dead_time_outer; p->fifo_dim = h->fifo_dim; p->fifo_slot = 0;
  p->adv_k = (real)(c.angle_advance * c.tau * c.motor_param[GEMB200_MP_P] * (sizeof(real) == 4 ? 1.0 / (2 * M_PI) : 1.0));
  p->inv_nsteps = (real)(1.0 / c.solver_nsteps);
  p->motor_kind = c.motor_kind;
  p->conv_kind[0] = c.converter_kind[0]; p->conv_kind[1] = c.converter_kind[1];
  p->load_kind = c.load_kind; p->solver_kind = c.solver_kind; p->nsteps = c.solver_nsteps;
  p->autoreset = c.autoreset;
  p->two_segment = h->two_segment;
  const double til0 = c.interlocking_time, til1 = c.interlocking_time1 < 0 ? c.interlocking_time : c.interlocking_time1;
  p->tau = (real)c.tau;
  p->til2[0] = (real)til0; p->til2[1] = (real)til1;
  p->tot2[0] = (real)(til0 / c.tau); p->tot2[1] = (real)(til1 / c.tau);
  p->lo_slot = til1 < til0 ? 1 : 0;
  p->promote = std::fabs(til1 - til0) - c.tau / 1000 > 0;  // t_hi - tau/1000 > t_start + t_lo (converters.py:273)
  const double hs[6] = {c.tau, til0, c.tau - til0, til1, c.tau - til1, std::fabs(til1 - til0)};
  for (int q = 0; q < 6; ++q) p->seg_len[q] = (real)hs[q];
  p->u_sup = (real)c.u_sup;
  {  // angle increment factors (see StepParams::kang) for every segment length of seg_len
    const double pp = c.motor_param[GEMB200_MP_P];
    const double unit = sizeof(real) == 4 ? 1.0 / (2 * M_PI) : 1.0;  // fp32 build keeps the angle in turns
    for (int sidx = 0; sidx < 6; ++sidx) {
      const double k_tot = pp * hs[sidx] * unit;
      const double k_sub = pp * (hs[sidx] / c.solver_nsteps) * (c.solver_kind == GEMB200_SOLVER_RK4 ? 1.0 / 6.0 : 1.0) * unit;
      const double ks[2] = {k_tot, k_sub};
      for (int m = 0; m < 2; ++m) {
        p->kang[m][sidx][0] = (real)ks[m];
        p->kang[m][sidx][1] = (real)(ks[m] - (double)p->kang[m][sidx][0]);
      }
    }
  }
  for (int j = 0; j < 20; ++j) p->k.c[j] = (real)dv.c[j];
  for (int j = 0; j < 4; ++j) p->k.tq[j] = (real)dv.tq[j];
  p->k.load_a = (real)c.load_param[GEMB200_LP_A]; p->k.load_b = (real)c.load_param[GEMB200_LP_B]; p->k.load_c = (real)c.load_param[GEMB200_LP_C];
  p->k.inv_j = (real)dv.inv_j; p->k.omega_lim = (real)dv.omega_lim; p->k.omega_lin = (real)dv.omega_lin;
  for (int j = 0; j < dm.n_state; ++j) {
    p->inv_lim[j] = c.limits[j] != 0.0 ? (real)(1.0 / c.limits[j]) : real(0);
    p->reset_obs[j] = (real)dv.reset_obs[j];
  }
  for (int j = 0; j < dm.nx; ++j) p->init_x[j] = (real)c.init_ode[j];
  if (dm.has_eps) {
    double e = c.init_ode[dm.nx];
    e = e - 2 * M_PI * std::rint(e / (2 * M_PI));
    if (e <= -M_PI) e += 2 * M_PI;
    const int eps_idx = dm.n_state - 2;  // [..., epsilon, u_sup]
    if (sizeof(real) == 4) {
      const double t = e / (2 * M_PI);
      p->init_ang[0] = (real)t; p->init_ang[1] = (real)(t - (double)p->init_ang[0]);
      p->eps_out_scale = (real)(2 * M_PI / c.limits[eps_idx]);
    } else {
      p->init_ang[0] = (real)e; p->init_ang[1] = real(0);
      p->eps_out_scale = (real)(1.0 / c.limits[eps_idx]);
    }
    p->inv_lim[eps_idx] = real(1);  // the angle entry is already normalised by eps_out_scale
  }
  p->init_random = c.init_random;
  p->init_im_valid = c.init_im_valid;
  for (int j = 0; j < 8; ++j) p->init_im[j] = (real)c.init_im[j];
  p->im_prev = static_cast<real*>(h->d_imprev);
  {  // truncated-normal states: CDF bounds prepared in double; the angle entry is converted to the stored unit like init_lo
    const int nst = dm.nx + (dm.has_eps ? 1 : 0);
    for (int j = 0; j < nst; ++j) {
      if (!c.init_random || !c.init_dist[j]) continue;
      const double unit = (j == dm.nx && sizeof(real) == 4) ? 1.0 / (2 * M_PI) : 1.0;
      p->init_mid[j] = std::isnan(c.init_mu[j]);  // mue = middle of the (possibly per-env) interval
      const double mu = p->init_mid[j] ? 0.5 * (c.init_hi[j] - c.init_lo[j]) + c.init_lo[j] : c.init_mu[j], sg = c.init_sigma[j];
      const double ca = 0.5 * std::erfc(-(c.init_lo[j] - mu) / sg * M_SQRT1_2), cb = 0.5 * std::erfc(-(c.init_hi[j] - mu) / sg * M_SQRT1_2);
      p->init_gauss = 1; p->init_dist[j] = 1;
      p->init_mu[j] = (real)(mu * unit); p->init_sigma[j] = (real)(sg * unit);
      p->init_ca[j] = (real)ca; p->init_cspan[j] = (real)(cb - ca);
    }
  }
  for (int j = 0; j < dm.nx; ++j) { p->init_lo[j] = (real)c.init_lo[j]; p->init_span[j] = (real)(c.init_hi[j] - c.init_lo[j]); }
  if (dm.has_eps) {
    const double unit = sizeof(real) == 4 ? 1.0 / (2 * M_PI) : 1.0;
    p->init_lo[dm.nx] = (real)(c.init_lo[dm.nx] * unit); p->init_span[dm.nx] = (real)((c.init_hi[dm.nx] - c.init_lo[dm.nx]) * unit);
  }
  for (int i = 0; i < c.n_constraints; ++i) {
    if (c.constraint_kind[i] == GEMB200_CONSTRAINT_SQUARED) {
      int cnt = 0;
      for (int j = 0; j < dm.n_obs; ++j) if ((c.constraint_mask[i] >> j) & 1u) p->sq_idx[p->n_sq][cnt++] = j;
      p->sq_cnt[p->n_sq++] = cnt;
    } else {
      for (int j = 0; j < dm.n_obs; ++j) {
        if (!((c.constraint_mask[i] >> j) & 1u)) continue;
        bool seen = false;
        for (int q = 0; q < p->n_lim; ++q) seen = seen || p->lim_idx[q] == j;
        if (!seen) p->lim_idx[p->n_lim++] = j;
      }
    }
  }
  // WeightedSumOfErrors: only non-zero weights become terms (weighted_sum_of_errors.py:128-129)
  int t = 0;
  for (int j = 0; j < dm.n_obs; ++j) {
    if (c.reward_weight[j] == 0.0) continue;
    int slot = -1;
    for (int r = 0; r < c.n_ref; ++r) if (c.ref_state[r] == j) slot = r;  // the last generator of a state wins (multiple_reference_generator.py:70-78)
    if (slot >= 0) {
      p->rwr_w[slot] = (real)c.reward_weight[j];
      p->rwr_inv_len[slot] = (real)(1.0 / c.state_length[j]);
      p->rwr_pow[slot] = (real)c.reward_power[j];
      p->rwr_pow1[slot] = c.reward_power[j] == 1.0;
      continue;
    }
    p->rw_state[t] = j;
    p->rw_w[t] = (real)c.reward_weight[j];
    p->rw_inv_len[t] = (real)(1.0 / c.state_length[j]);
    p->rw_pow[t] = (real)c.reward_power[j];
    p->rw_pow1[t] = c.reward_power[j] == 1.0;
    ++t;
  }
  p->n_rw = t;
  for (int r = 0; r < kMaxRef; ++r) if (p->rwr_w[r] == real(0)) p->rwr_pow1[r] = 1;  // unused slots: no pow()
  p->bias = (real)c.reward_bias; p->viol_reward = (real)c.violation_reward;
  // PLAIN shape (step_kernel): decided here once; GEMB200_NO_PLAIN=1 in the environment forces the general instantiation (A/B runs)
  {
    bool plain =
                 c.load_kind != GEMB200_LOAD_EXT_SPEED && c.supply_kind == GEMB200_SUPPLY_IDEAL && (c.finite || (c.interlocking_time == 0.0 && !(c.interlocking_time1 > 0.0))) && c.dead_time_steps == 0 && !c.action_dq && c.n_state_ops == 0 &&
                 c.converter_kind[0] != GEMB200_CONV_1QC && c.converter_kind[1] != GEMB200_CONV_1QC &&
                 p->n_rw == 0 && p->n_lim <= 2 && p->n_sq <= 1 && (p->n_sq == 0 || p->sq_cnt[0] == 2);
    for (int r = 0; r < c.n_ref; ++r) plain = plain && c.ref_kind[r] == GEMB200_REF_WIENER && p->rwr_pow1[r] && c.ref_sw_count[r] <= 1;
    // PLAIN monitor, branch-free: word offsets of the (at most) two limit-checked states and the two states of the squared constraint in the
    // staged row; an unused check compares entry 0 with +inf
    const real inf = std::numeric_limits<real>::infinity();
    for (int q = 0; q < 2; ++q) { p->mon_off[q] = p->n_lim > q ? p->lim_idx[q] * (int)sizeof(real) : 0; p->mon_thr[q] = p->n_lim > q ? real(1) : inf; }
    for (int q = 0; q < 2; ++q) p->mon_off[2 + q] = (p->n_sq > 0 && p->sq_cnt[0] == 2) ? p->sq_idx[0][q] * (int)sizeof(real) : 0;
    p->mon_thr[2] = p->n_sq > 0 ? real(1) : inf;
    const char* off = std::getenv("GEMB200_NO_PLAIN");
    p->plain = plain && !(off && off[0] == '1');
    const_cast<gemb200_handle*>(h)->plain_shape = p->plain;
  }
  {  // L2 prefetch distance: one wave of resident threads (SMs x blocks/SM x block size); GEMB200_PF_DIST overrides (0 = off)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    p->pf_dist = sms * 4 * GEMB200_BLOCK;  // measured optimum 0.25-1 wave, flat (profiles/r01_variants.md)
    if (const char* e = std::getenv("GEMB200_PF_DIST")) p->pf_dist = std::atoi(e);
  }
  p->ext_tab = static_cast<const real*>(h->d_ext);
  p->ext_len = c.ext_speed_len;
  p->ext_inv_tau = c.load_kind == GEMB200_LOAD_EXT_SPEED ? (real)(1.0 / c.load_param[GEMB200_LP_TAU_LOAD]) : real(0);
  p->kenv = h->d_kenv;
  p->supply_kind = c.supply_kind;
  p->sup = static_cast<real*>(h->d_sup);
  p->sup_k1 = c.supply_kind == GEMB200_SUPPLY_RC ? (real)(c.tau / (c.supply_param[0] * c.supply_param[1])) : real(0);
  p->sup_k2 = (real)c.supply_param[0];
  p->sup_phase = h->d_supph;
  if (c.supply_kind == GEMB200_SUPPLY_AC1) {
    const double unit = sizeof(real) == 4 ? 1.0 / (2 * M_PI) : 1.0;  // fp32 build: turns as double-float
    const double kph = 2 * M_PI * c.supply_param[0] * c.tau * unit;
    double ph0 = c.supply_param[1];
    ph0 = ph0 - 2 * M_PI * std::rint(ph0 / (2 * M_PI));
    ph0 *= unit;
    p->sup_amp = (real)(std::sqrt(2.0) * c.u_sup);
    p->sup_kph[0] = (real)kph; p->sup_kph[1] = sizeof(real) == 4 ? (real)(kph - (double)p->sup_kph[0]) : real(0);
    p->sup_ph0[0] = (real)ph0; p->sup_ph0[1] = sizeof(real) == 4 ? (real)(ph0 - (double)p->sup_ph0[0]) : real(0);
    p->sup_fixed = c.supply_param[2] != 0.0;
  }
  for (int j = 0; j < dm.n_state; ++j) p->reset_obs_du[j] = (real)dv.reset_obs_du[j];
  p->n_sops = c.n_state_ops;
  p->n_obs = dm.n_obs;
  p->row_stride = h->row_stride;
  p->obsv = static_cast<real*>(h->d_obsv);
  for (int k = 0; k < c.n_state_ops; ++k) {
    p->sop_kind[k] = c.sop_kind[k];
    p->sop_mask[k] = c.sop_mask[k];
    for (int q = 0; q < 4; ++q) p->sop_idx[k][q] = c.sop_idx[k][q];
    for (int q = 0; q < 8; ++q) p->sop_param[k][q] = (real)c.sop_param[k][q];
  }
  p->n_ref = c.n_ref;
  p->swst = h->d_swst;
  for (int r = 0; r < c.n_ref; ++r) {
    p->sw_count[r] = c.ref_sw_count[r] > 1 ? c.ref_sw_count[r] : 0;
    p->sw_first[r] = c.ref_sw_first[r];
    p->sw_len_lo[r] = c.ref_sw_len_lo[r]; p->sw_len_span[r] = c.ref_sw_len_hi[r] - c.ref_sw_len_lo[r];
  }
  for (int r = 0; r < GEMB200_MAX_REF_ENTRIES; ++r) p->sw_cdf[r] = (real)c.ref_sw_cdf[r];
  p->any_wiener = h->any_wiener;
  p->ref_tau = (real)c.tau;
  for (int r = 0; r < GEMB200_MAX_REF_ENTRIES; ++r) {  // all parameter entries (switched sub-generators live beyond n_ref)
    p->ref_kind[r] = c.ref_kind[r]; p->ref_state[r] = c.ref_state[r];
    p->ref_const[r] = (real)c.ref_value[r];
    p->ref_lo[r] = (real)c.ref_margin_lo[r]; p->ref_hi[r] = (real)c.ref_margin_hi[r];
    p->ref_init_lo[r] = (real)c.ref_init_lo[r]; p->ref_init_span[r] = (real)(c.ref_init_hi[r] - c.ref_init_lo[r]);
    p->ref_amp_lo[r] = (real)c.ref_amp_lo[r]; p->ref_amp_span[r] = (real)(c.ref_amp_hi[r] - c.ref_amp_lo[r]);
    p->ref_freq_lo[r] = (real)c.ref_freq_lo[r]; p->ref_freq_span[r] = (real)(c.ref_freq_hi[r] - c.ref_freq_lo[r]);
    p->ref_off_lo[r] = (real)c.ref_off_lo[r]; p->ref_off_hi[r] = (real)c.ref_off_hi[r];
    if (c.ref_kind[r] == GEMB200_REF_WIENER || c.ref_kind[r] == GEMB200_REF_LAPLACE) {
      p->ref_lsig_lo[r] = (real)std::log10(c.ref_sigma_lo[r]);
      p->ref_lsig_span[r] = (real)(std::log10(c.ref_sigma_hi[r]) - std::log10(c.ref_sigma_lo[r]));
    }
    p->ref_len_lo[r] = c.ref_len_lo[r]; p->ref_len_span[r] = c.ref_len_hi[r] - c.ref_len_lo[r];
  }
}

// ----------------------------------------------------------------------------------------------------------------
// kernel dispatch (the instantiations live in gemb200_step_tu.cu, one TU per family x real)
// ----------------------------------------------------------------------------------------------------------------
// GEMB200_ONLY_FAM=<family>: experiment builds (tools/build_variants.py --only) that carry the fp32 kernels of ONE motor family; every
// other configuration fails with cudaErrorInvalidValue instead of leaving unresolved symbols.  Never defined in the product build.
template <typename real>
static cudaError_t launch_step(int fam, bool finite, int nref, const StepParams<real>& p, cudaStream_t st) {
#ifdef GEMB200_ONLY_FAM
  if constexpr (std::is_same<real, float>::value) { if (fam == GEMB200_ONLY_FAM) return launch_step_f<GEMB200_ONLY_FAM, real>(finite, nref, p, st); }
  return cudaErrorInvalidValue;
#else
  switch (fam) {
    case kDC1: return launch_step_f<kDC1, real>(finite, nref, p, st);
    case kDC2: return launch_step_f<kDC2, real>(finite, nref, p, st);
    case kSYNC: return launch_step_f<kSYNC, real>(finite, nref, p, st);
    case kEESM: return launch_step_f<kEESM, real>(finite, nref, p, st);
    case kSCIM: return launch_step_f<kSCIM, real>(finite, nref, p, st);
    case kDFIM: return launch_step_f<kDFIM, real>(finite, nref, p, st);
  }
  return cudaErrorInvalidValue;
#endif
}
template <typename real>
static cudaError_t launch_reset(int fam, int nref, const StepParams<real>& p, cudaStream_t st) {
#ifdef GEMB200_ONLY_FAM
  if constexpr (std::is_same<real, float>::value) { if (fam == GEMB200_ONLY_FAM) return launch_reset_f<GEMB200_ONLY_FAM, real>(nref, p, st); }
  return cudaErrorInvalidValue;
#else
  switch (fam) {
    case kDC1: return launch_reset_f<kDC1, real>(nref, p, st);
    case kDC2: return launch_reset_f<kDC2, real>(nref, p, st);
    case kSYNC: return launch_reset_f<kSYNC, real>(nref, p, st);
    case kEESM: return launch_reset_f<kEESM, real>(nref, p, st);
    case kSCIM: return launch_reset_f<kSCIM, real>(nref, p, st);
    case kDFIM: return launch_reset_f<kDFIM, real>(nref, p, st);
  }
  return cudaErrorInvalidValue;
#endif
}

template <typename real>
static void set_roll_strides(const gemb200_handle* h, StepParams<real>& p) {
  const int64_t n = h->cfg.n_envs;
  p.out_has = (p.obs ? 1 : 0) | ((p.ref_out && h->n_ref > 0) ? 2 : 0) | (p.reward ? 4 : 0) | (p.term ? 8 : 0);
  p.roll_act_inc = n * h->n_act * (int64_t)(h->cfg.finite ? sizeof(int32_t) : sizeof(real));
  p.roll_obs_inc = p.obs ? n * h->n_obs * (int64_t)sizeof(real) : 0;
  p.roll_ref_inc = p.ref_out ? n * h->n_ref * (int64_t)sizeof(real) : 0;
  p.roll_rew_inc = p.reward ? n * (int64_t)sizeof(real) : 0;
  p.roll_term_inc = p.term ? n : 0;
}

// ---- device-resident clock: the call id of the NEXT call, the step count and the dead-time ring position live in device memory and are
// advanced by a one-thread kernel behind every launch, so that a launch depends on nothing the host changes between calls (CUDA graphs)
__global__ void clock_tick_kernel(uint32_t* c, uint32_t d_call, uint32_t d_step, uint32_t dead_steps) {
  const uint64_t g = (((uint64_t)c[1] << 32) | c[0]) + d_call, s = (((uint64_t)c[4] << 32) | c[2]) + d_step;
  c[0] = (uint32_t)g; c[1] = (uint32_t)(g >> 32); c[2] = (uint32_t)s; c[4] = (uint32_t)(s >> 32);
  c[3] = dead_steps ? (uint32_t)(s % dead_steps) : 0u;
}
static int push_clock(gemb200_handle* h, cudaStream_t st) {  // host counters -> device (stream-ordered; the source is a by-value kernel argument)
  const uint64_t g1 = h->gstep + 1, s = h->n_steps;
  const uint32_t dead = (uint32_t)h->cfg.dead_time_steps;
  const uint32_t v[8] = {(uint32_t)g1, (uint32_t)(g1 >> 32), (uint32_t)s, dead ? (uint32_t)(s % dead) : 0u, (uint32_t)(s >> 32), 0u, 0u, 0u};
  CUDA_TRY(cudaMemcpyAsync(h->d_clock, v, sizeof(v), cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
  return GEMB200_OK;
}
static int pull_clock(gemb200_handle* h, cudaStream_t st) {  // device -> host counters (synchronises the stream)
  uint32_t v[8];
  CUDA_TRY(cudaMemcpyAsync(v, h->d_clock, sizeof(v), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  h->gstep = ((((uint64_t)v[1]) << 32) | v[0]) - 1;
  h->n_steps = (((uint64_t)v[4]) << 32) | v[2];
  return GEMB200_OK;
}
static int tick_clock(gemb200_handle* h, uint32_t d_call, uint32_t d_step, cudaStream_t st) {
  clock_tick_kernel<<<1, 1, 0, st>>>(h->d_clock, d_call, d_step, (uint32_t)h->cfg.dead_time_steps);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GEMB200_OK;
}

// One launch over envs [begin, end) (end < 0: all).  new_call: this launch starts a new API call (fresh RNG call ids);
// the chunks of one pipelined host step share them.  roll > 0: `roll` fused steps (rollout_kernel) whose call ids, step clock and
// dead-time ring positions are exactly those of `roll` consecutive single-step calls; outputs every `every` steps (0: last only).
static int do_step(gemb200_handle* h, const void* action, void* obs, void* ref, void* rew, uint8_t* term, cudaStream_t st,
                   int begin = 0, int end = -1, bool new_call = true, int roll = 0, int every = 0) {
  if (!action) return fail(GEMB200_E_INVALID, "action is NULL");
  const uint64_t ksteps = roll > 0 ? (uint64_t)roll : 1;
  const bool dev_clock = h->dev_clock;
  if (dev_clock && (!new_call || begin != 0 || (end >= 0 && end != h->cfg.n_envs)))
    return fail(GEMB200_E_INVALID, "the host-buffer step is not available while the device-resident clock is enabled");
  if (new_call && !dev_clock) { h->gstep += ksteps; h->n_steps += ksteps; }
  const uint64_t g0 = h->gstep - (ksteps - 1), n0 = h->n_steps - (ksteps - 1);  // call id / step count of the FIRST step of this launch
  if (end < 0) end = h->cfg.n_envs;
  const int fifo_slot = h->cfg.dead_time_steps > 0 ? (int)((n0 - 1) % (uint64_t)h->cfg.dead_time_steps) : 0;
  cudaError_t e;
  if (h->cfg.dtype == GEMB200_F32) {
    StepParams<float>& p = h->pf;
    p.env_begin = begin; p.env_end = end; p.fifo_slot = fifo_slot; p.kstep = dev_clock ? 1u : (uint32_t)n0;
    p.gstep_lo = (uint32_t)g0; p.gstep_hi = (uint32_t)(g0 >> 32); p.clock_dev = dev_clock ? h->d_clock : nullptr;
    p.roll_steps = roll; p.record_every = every;
    p.action = action; p.obs = (float*)obs; p.ref_out = (float*)ref; p.reward = (float*)rew; p.term = term;
    set_roll_strides(h, p);
    e = launch_step<float>(h->fam, h->cfg.finite != 0, h->n_ref, p, st);
  } else {
    StepParams<double>& p = h->pd;
    p.env_begin = begin; p.env_end = end; p.fifo_slot = fifo_slot; p.kstep = dev_clock ? 1u : (uint32_t)n0;
    p.gstep_lo = (uint32_t)g0; p.gstep_hi = (uint32_t)(g0 >> 32); p.clock_dev = dev_clock ? h->d_clock : nullptr;
    p.roll_steps = roll; p.record_every = every;
    p.action = action; p.obs = (double*)obs; p.ref_out = (double*)ref; p.reward = (double*)rew; p.term = term;
    set_roll_strides(h, p);
    e = launch_step<double>(h->fam, h->cfg.finite != 0, h->n_ref, p, st);
  }
  if (e != cudaSuccess) return fail(GEMB200_E_CUDA, std::string(roll > 0 ? "rollout launch: " : "step launch: ") + cudaGetErrorString(e));
  h->launches += 1;
  // (advancing the clock from the step kernel itself — its last block to finish, one atomic per block — was measured and is SLOWER in a graph
  // than this separate one-thread node: 5.5 vs 4.7 us per captured step at N = 65 536, where the grid has 2048 small blocks)
  if (dev_clock) return tick_clock(h, (uint32_t)ksteps, (uint32_t)ksteps, st);
  return GEMB200_OK;
}

static int do_reset(gemb200_handle* h, const uint8_t* mask, void* obs, void* ref, cudaStream_t st) {
  const bool dev_clock = h->dev_clock;
  if (!dev_clock) h->gstep += 1;
  h->pf.clock_dev = h->pd.clock_dev = dev_clock ? h->d_clock : nullptr;
  cudaError_t e;
  if (h->cfg.dtype == GEMB200_F32) {
    StepParams<float>& p = h->pf;
    p.gstep_lo = (uint32_t)h->gstep; p.gstep_hi = (uint32_t)(h->gstep >> 32);
    p.reset_mask = mask; p.obs = (float*)obs; p.ref_out = (float*)ref; p.kstep = dev_clock ? 0u : (uint32_t)h->n_steps;
    e = launch_reset<float>(h->fam, h->n_ref, p, st);
    p.reset_mask = nullptr;
  } else {
    StepParams<double>& p = h->pd;
    p.gstep_lo = (uint32_t)h->gstep; p.gstep_hi = (uint32_t)(h->gstep >> 32);
    p.reset_mask = mask; p.obs = (double*)obs; p.ref_out = (double*)ref; p.kstep = dev_clock ? 0u : (uint32_t)h->n_steps;
    e = launch_reset<double>(h->fam, h->n_ref, p, st);
    p.reset_mask = nullptr;
  }
  if (e != cudaSuccess) return fail(GEMB200_E_CUDA, std::string("reset launch: ") + cudaGetErrorString(e));
  h->launches += 1;
  if (dev_clock) return tick_clock(h, 1u, 0u, st);
  return GEMB200_OK;
}

// induction motors: before the first reset the "previous" initial currents are the constant ones (_initial_states of a fresh motor)
static int fill_imprev(gemb200_handle* h, cudaStream_t st) {
  if (!h->d_imprev) return GEMB200_OK;
  const size_t n = (size_t)h->cfg.n_envs;
  if (h->cfg.dtype == GEMB200_F32) {
    std::vector<float> v(2 * n);
    for (size_t q = 0; q < n; ++q) { v[q] = (float)h->cfg.init_ode[1]; v[n + q] = (float)h->cfg.init_ode[2]; }
    CUDA_TRY(cudaMemcpyAsync(h->d_imprev, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  } else {
    std::vector<double> v(2 * n);
    for (size_t q = 0; q < n; ++q) { v[q] = h->cfg.init_ode[1]; v[n + q] = h->cfg.init_ode[2]; }
    CUDA_TRY(cudaMemcpyAsync(h->d_imprev, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return GEMB200_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ----------------------------------------------------------------------------------------------------------------
// C-ABI
// ----------------------------------------------------------------------------------------------------------------
extern "C" {

int gemb200_version(void) { return GEMB200_ABI_VERSION; }
const char* gemb200_last_error(void) { return g_last_error.c_str(); }

int gemb200_config_init(gemb200_config* cfg) {
  if (!cfg) return fail(GEMB200_E_INVALID, "config is NULL");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(gemb200_config);
  cfg->abi_version = GEMB200_ABI_VERSION;
  cfg->n_envs = 1;
  cfg->solver_kind = GEMB200_SOLVER_RK4;
  cfg->solver_nsteps = 1;
  cfg->tau = 1e-4;
  cfg->load_param[GEMB200_LP_TAU_DECAY] = 1e-3;
  cfg->interlocking_time1 = -1.0;
  for (int i = 0; i < GEMB200_MAX_STATE; ++i) { cfg->limits[i] = 1.0; cfg->state_length[i] = 2.0; cfg->reward_power[i] = 1.0; }
  for (int r = 0; r < GEMB200_MAX_REF_ENTRIES; ++r) {
    cfg->ref_len_lo[r] = 500; cfg->ref_len_hi[r] = 2000;
    cfg->ref_sigma_lo[r] = 1e-3; cfg->ref_sigma_hi[r] = 1e-1;
    cfg->ref_margin_lo[r] = -1; cfg->ref_margin_hi[r] = 1; cfg->ref_init_lo[r] = -1; cfg->ref_init_hi[r] = 1;
  }
  return GEMB200_OK;
}

int gemb200_query_dims(const gemb200_config* cfg, int32_t* n_state, int32_t* n_ode, int32_t* n_act, int32_t* n_ref) {
  if (!cfg) return fail(GEMB200_E_INVALID, "config is NULL");
  Dims d;
  int rc = derive_dims(cfg, &d);
  if (rc) return rc;
  if (n_state) *n_state = d.n_obs;
  if (n_ode) *n_ode = d.n_ode;
  if (n_act) *n_act = d.n_act;
  if (n_ref) *n_ref = cfg->n_ref;
  return GEMB200_OK;
}

int gemb200_create(const gemb200_config* cfg, gemb200_handle** out) {
  if (!out) return fail(GEMB200_E_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(GEMB200_E_INVALID, "device ordinal out of range");
  DeviceGuard guard(cfg->device);
  gemb200_handle* h = new (std::nothrow) gemb200_handle();
  if (!h) return fail(GEMB200_E_NOMEM, "out of host memory");
  h->cfg = *cfg;
  Dims d;
  derive_dims(cfg, &d);
  h->fam = d.fam; h->n_state = d.n_state; h->n_ode = d.n_ode; h->n_act = d.n_act; h->nx = d.nx; h->has_eps = d.has_eps;
  h->n_obs = d.n_obs;
  h->row_stride = cfg->n_state_ops > 0 ? (d.n_obs | 1) : (d.fam == kEESM ? 17 : (d.fam == kDFIM ? 25 : d.n_state));  // Fam<>::PAD without wrappers
  h->n_ref = cfg->n_ref;
  h->rsz = cfg->dtype == GEMB200_F32 ? 4 : 8;
  h->two_segment = cfg->finite && (cfg->interlocking_time > 0 || cfg->interlocking_time1 > 0);
  for (int r = 0; r < GEMB200_MAX_REF_ENTRIES; ++r) {  // any generator that advances by itself (Wiener, Laplace, periodic), incl. switched subs
    bool used = r < cfg->n_ref;
    for (int q = 0; q < cfg->n_ref; ++q) used = used || (cfg->ref_sw_count[q] > 1 && r >= cfg->ref_sw_first[q] && r < cfg->ref_sw_first[q] + cfg->ref_sw_count[q]);
    if (used) h->any_wiener = h->any_wiener || cfg->ref_kind[r] == GEMB200_REF_WIENER || cfg->ref_kind[r] >= GEMB200_REF_LAPLACE;
    if (r < cfg->n_ref && cfg->ref_sw_count[r] > 1) h->any_switched = true;
  }
  const size_t n = (size_t)cfg->n_envs;
#define ALLOC(ptr, bytes)                                                                                     \
  do {                                                                                                        \
    cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                                     \
    if (e_ != cudaSuccess) { gemb200_destroy(h); return fail(e_ == cudaErrorMemoryAllocation ? GEMB200_E_NOMEM : GEMB200_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); } \
    cudaMemset((ptr), 0, (bytes));                                                                            \
  } while (0)
  h->NH = hot_words(d.nx, cfg->n_ref); h->NC = cold_words(d.nx, cfg->n_ref);
  ALLOC(h->d_st, n * (h->NH > 0 ? h->NH : 1) * h->rsz);
  ALLOC(h->d_stc, n * h->NC * h->rsz);
  if (cfg->dead_time_steps > 0) {
    // queue width: caller-side actions when the dead time wraps the dq transformation (or there is none), else abc(+e)
    const int inner = cfg->finite ? d.n_act : (d.fam == kDFIM ? 6 : (d.fam == kEESM ? 4 : (d.fam >= kSYNC ? 3 : d.n_act)));
    h->fifo_dim = (cfg->action_dq && !cfg->dead_time_outer) ? inner : d.n_act;
    ALLOC(h->d_fifo, n * cfg->dead_time_steps * h->fifo_dim * h->rsz);
  }
  if (d.has_eps) ALLOC(h->d_eps, n * sizeof(double));
  if (h->two_segment || (cfg->finite && cfg->supply_kind == GEMB200_SUPPLY_RC)) ALLOC(h->d_sw, n * sizeof(uint16_t));
  if (cfg->supply_kind == GEMB200_SUPPLY_RC) ALLOC(h->d_sup, n * 2 * h->rsz);
  if (cfg->supply_kind == GEMB200_SUPPLY_AC1) ALLOC(h->d_supph, n * sizeof(double));
  if (cfg->load_kind == GEMB200_LOAD_EXT_SPEED) {
    ALLOC(h->d_kenv, n * sizeof(uint32_t));
    ALLOC(h->d_ext, (size_t)cfg->ext_speed_len * h->rsz);
    if (cfg->dtype == GEMB200_F32) {
      std::vector<float> tmp(cfg->ext_speed_table, cfg->ext_speed_table + cfg->ext_speed_len);
      cudaMemcpy(h->d_ext, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice);
    } else {
      cudaMemcpy(h->d_ext, cfg->ext_speed_table, (size_t)cfg->ext_speed_len * sizeof(double), cudaMemcpyHostToDevice);
    }
    {
      uint64_t x = 1469598103934665603ull;
      const unsigned char* tb = reinterpret_cast<const unsigned char*>(cfg->ext_speed_table);
      for (size_t q = 0; q < (size_t)cfg->ext_speed_len * sizeof(double); ++q) { x ^= tb[q]; x *= 1099511628211ull; }
      h->ext_hash = x;
    }
    h->cfg.ext_speed_table = nullptr;  // the caller's buffer is not referenced after create
  }
  if (h->any_switched) ALLOC(h->d_swst, n * 2 * cfg->n_ref * sizeof(uint32_t));
  if (d.has_observer) ALLOC(h->d_obsv, n * 4 * h->rsz);
  if (cfg->init_im_valid) ALLOC(h->d_imprev, n * 2 * h->rsz);
#undef ALLOC
  Derived dv;
  derive_model(cfg, d, &dv);
  {  // same derivation one volt higher: the difference is the u_sup-proportional part of the reset observation
    gemb200_config c1 = *cfg;
    c1.u_sup += 1.0;
    Derived dv1;
    derive_model(&c1, d, &dv1);
    for (int j = 0; j < d.n_state; ++j) dv.reset_obs_du[j] = dv1.reset_obs[j] - dv.reset_obs[j];
  }
  fill_params<float>(h, d, dv, &h->pf);
  fill_params<double>(h, d, dv, &h->pd);
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  rc = fill_imprev(h, nullptr);
  if (rc) { gemb200_destroy(h); return rc; }
  rc = do_reset(h, nullptr, nullptr, nullptr, nullptr);
  if (rc) { gemb200_destroy(h); return rc; }
  cudaError_t e = cudaStreamSynchronize(nullptr);
  if (e != cudaSuccess) { gemb200_destroy(h); return fail(GEMB200_E_CUDA, std::string("initial reset: ") + cudaGetErrorString(e)); }
  *out = h;
  return GEMB200_OK;
}

int gemb200_destroy(gemb200_handle* h) {
  if (!h) return GEMB200_OK;
  DeviceGuard guard(h->cfg.device);
  cudaFree(h->d_st); cudaFree(h->d_stc); cudaFree(h->d_eps); cudaFree(h->d_sw); cudaFree(h->d_fifo); cudaFree(h->d_obsv); cudaFree(h->d_sup); cudaFree(h->d_supph); cudaFree(h->d_swst); cudaFree(h->d_ext); cudaFree(h->d_kenv); cudaFree(h->d_imprev); cudaFree(h->d_envp); cudaFree(h->d_clock);
  cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_ref); cudaFree(h->d_rew); cudaFree(h->d_term); cudaFree(h->d_mask);
  if (h->hstream) cudaStreamDestroy(h->hstream);
  for (int k = 0; k < 3; ++k) if (h->hpipe[k]) cudaStreamDestroy(h->hpipe[k]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return GEMB200_OK;
}

int gemb200_reset(gemb200_handle* h, const uint8_t* reset_mask, void* obs_out, void* ref_out, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  return do_reset(h, reset_mask, obs_out, ref_out, (cudaStream_t)stream);
}

int gemb200_step(gemb200_handle* h, const void* action, void* obs_out, void* ref_out, void* reward_out, uint8_t* terminated_out, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  return do_step(h, action, obs_out, ref_out, reward_out, terminated_out, (cudaStream_t)stream);
}

int gemb200_rollout_record(gemb200_handle* h, const void* actions, int32_t n_steps, int32_t record_every, void* obs_out, void* ref_out,
                           void* reward_out, uint8_t* terminated_out, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  if (n_steps < 1 || n_steps > (1 << 24)) return fail(GEMB200_E_INVALID, "n_steps must be in [1, 2^24]");
  if (record_every < 0 || record_every > n_steps) return fail(GEMB200_E_INVALID, "record_every must be in [0, n_steps]");
  DeviceGuard guard(h->cfg.device);
  return do_step(h, actions, obs_out, ref_out, reward_out, terminated_out, (cudaStream_t)stream, 0, -1, true, n_steps, record_every);
}

int gemb200_rollout(gemb200_handle* h, const void* actions, int32_t n_steps, void* obs_out, void* ref_out, void* reward_out,
                    uint8_t* terminated_out, void* stream) {
  return gemb200_rollout_record(h, actions, n_steps, 0, obs_out, ref_out, reward_out, terminated_out, stream);
}

// Domain randomisation (the batched counterpart of constructing N reference envs with N motor_parameter / load_parameter
// dicts): every env gets its own model coefficients, derived here exactly like the shared ones (the *_update_model methods,
// mechanical_load.py:188-193) from ITS physical parameters.  Limits, nominal values, reward and reference settings stay those of the
// handle's configuration.  NULL motor_param: back to the shared coefficients.
int gemb200_set_env_params(gemb200_handle* h, const double* motor_param, const double* load_param) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(cudaDeviceSynchronize());
  if (h->cfg.layout != GEMB200_LAYOUT_AOS && (motor_param || load_param))
    return fail(GEMB200_E_INVALID, "per-env parameter blocks need the row-per-env (AoS) I/O layout");
  if (!motor_param && !load_param) {
    h->pf.envp = nullptr; h->pd.envp = nullptr;
    h->pf.plain = h->pf.n_dst == 0 ? h->plain_shape : 0; h->pd.plain = h->pf.plain;
    return GEMB200_OK;
  }
  const size_t n = (size_t)h->cfg.n_envs;
  Dims d;
  derive_dims(&h->cfg, &d);
  std::vector<double> tab((size_t)kCoefWords * n);
  gemb200_config c = h->cfg;
  for (size_t i = 0; i < n; ++i) {
    if (motor_param) std::memcpy(c.motor_param, motor_param + i * GEMB200_MAX_MOTOR_PARAM, sizeof(c.motor_param));
    if (load_param) std::memcpy(c.load_param, load_param + i * 8, sizeof(c.load_param));
    if (c.load_kind == GEMB200_LOAD_POLY_STATIC && !(c.load_param[GEMB200_LP_J_LOAD] + c.motor_param[GEMB200_MP_J_ROTOR] > 0))
      return fail(GEMB200_E_INVALID, "per-env parameters: total inertia must be positive for every env");
    Derived dv;
    derive_model(&c, d, &dv);
    for (int w = 0; w < 20; ++w) tab[(size_t)w * n + i] = dv.c[w];
    for (int w = 0; w < 4; ++w) tab[(size_t)(20 + w) * n + i] = dv.tq[w];
    tab[(size_t)24 * n + i] = c.load_param[GEMB200_LP_A]; tab[(size_t)25 * n + i] = c.load_param[GEMB200_LP_B]; tab[(size_t)26 * n + i] = c.load_param[GEMB200_LP_C];
    tab[(size_t)27 * n + i] = dv.inv_j; tab[(size_t)28 * n + i] = dv.omega_lim; tab[(size_t)29 * n + i] = dv.omega_lin;
  }
  for (double v : tab) if (!std::isfinite(v)) return fail(GEMB200_E_INVALID, "per-env parameters: a derived model coefficient is not finite (zero inductance?)");
  if (!h->d_envp) CUDA_TRY(cudaMalloc(&h->d_envp, tab.size() * h->rsz));
  if (h->cfg.dtype == GEMB200_F32) {
    std::vector<float> tf(tab.begin(), tab.end());
    CUDA_TRY(cudaMemcpy(h->d_envp, tf.data(), tf.size() * sizeof(float), cudaMemcpyHostToDevice));
  } else {
    CUDA_TRY(cudaMemcpy(h->d_envp, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  h->pf.envp = static_cast<const float*>(h->d_envp); h->pd.envp = static_cast<const double*>(h->d_envp);
  h->pf.plain = 0; h->pd.plain = 0;  // the PLAIN instantiations read the shared constant-bank coefficients
  return GEMB200_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// Fused aggregated return over NVLink (the sharded layout's ONE collective, done by the step kernel itself)
// ----------------------------------------------------------------------------------------------------------------
// Flags: one uint32 per (slot, source rank), written by the source with a system-scope store after its step kernel has finished (every
// thread of the step kernel fences its peer stores at system scope before it exits), polled by the owner.
__global__ void peer_signal_kernel(uint32_t* const* flags, int n, uint32_t value) {
  const int d = threadIdx.x;
  if (d < n) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(flags[d]) = value;
    __threadfence_system();
  }
}
// spins until all n flags are >= value (wrap-safe signed distance); gives up after ~4e9 cycles and raises *err so that a lost peer cannot hang the GPU
__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t value, int* err) {
  const int s = threadIdx.x;
  if (s < n) {
    const long long t0 = clock64();
    while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(flags + s) - value) < 0) {
      if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1 + s); break; }
      __nanosleep(200);
    }
    __threadfence_system();
  }
}

int gemb200_peer_buffer_alloc(int32_t device, int64_t bytes, void** dev_ptr, void* ipc_handle64) {
  if (!dev_ptr || !ipc_handle64 || bytes <= 0) return fail(GEMB200_E_INVALID, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard guard(device);
  CUDA_TRY(cudaMalloc(dev_ptr, (size_t)bytes));
  CUDA_TRY(cudaMemset(*dev_ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t hd;
  CUDA_TRY(cudaIpcGetMemHandle(&hd, *dev_ptr));
  std::memcpy(ipc_handle64, &hd, sizeof(hd));
  CUDA_TRY(cudaDeviceSynchronize());
  return GEMB200_OK;
}
int gemb200_peer_buffer_open(int32_t device, const void* ipc_handle64, void** dev_ptr) {
  if (!dev_ptr || !ipc_handle64) return fail(GEMB200_E_INVALID, "bad argument");
  DeviceGuard guard(device);  // the CONSUMER's device is current: the mapping lives in this context and peer access is enabled lazily
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, ipc_handle64, sizeof(hd));
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return GEMB200_OK;
}
int gemb200_peer_buffer_close(int32_t device, void* dev_ptr) {
  DeviceGuard guard(device);
  CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return GEMB200_OK;
}
int gemb200_peer_buffer_free(int32_t device, void* dev_ptr) {
  DeviceGuard guard(device);
  CUDA_TRY(cudaFree(dev_ptr));
  return GEMB200_OK;
}
int gemb200_bind_peers(gemb200_handle* h, int32_t n_dst, const int64_t* dst_delta) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  if (n_dst < 0 || n_dst > 8 || (n_dst > 0 && !dst_delta)) return fail(GEMB200_E_INVALID, "n_dst must be in [0, 8]");
  if (n_dst > 0 && h->cfg.layout != GEMB200_LAYOUT_AOS) return fail(GEMB200_E_INVALID, "peer destinations need the row-per-env (AoS) layout");
  h->pf.n_dst = n_dst; h->pd.n_dst = n_dst;
  for (int d = 0; d < n_dst; ++d) { h->pf.dst_delta[d] = dst_delta[d]; h->pd.dst_delta[d] = dst_delta[d]; }
  // the PLAIN instantiations store to the caller's tensors only
  h->pf.plain = (n_dst == 0 && !h->pf.envp) ? h->plain_shape : 0;
  h->pd.plain = h->pf.plain;
  return GEMB200_OK;
}
int gemb200_peer_signal(gemb200_handle* h, int32_t n_dst, uint32_t* const* flag_ptrs_dev, uint32_t value, void* stream) {
  if (!h || n_dst < 1 || n_dst > 32 || !flag_ptrs_dev) return fail(GEMB200_E_INVALID, "bad argument");
  DeviceGuard guard(h->cfg.device);
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_ptrs_dev, n_dst, value);
  CUDA_TRY(cudaGetLastError());
  return GEMB200_OK;
}
int gemb200_peer_wait(gemb200_handle* h, int32_t n_src, const uint32_t* flags_dev, uint32_t value, int32_t* err_dev, void* stream) {
  if (!h || n_src < 1 || n_src > 32 || !flags_dev || !err_dev) return fail(GEMB200_E_INVALID, "bad argument");
  DeviceGuard guard(h->cfg.device);
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_dev, n_src, value, err_dev);
  CUDA_TRY(cudaGetLastError());
  return GEMB200_OK;
}

static int ensure_host_buffers(gemb200_handle* h) {
  if (h->hstream) return GEMB200_OK;
  const size_t n = (size_t)h->cfg.n_envs;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
  for (int k = 0; k < 3; ++k) CUDA_TRY(cudaStreamCreateWithFlags(&h->hpipe[k], cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&h->d_act, n * h->n_act * (h->cfg.finite ? sizeof(int32_t) : h->rsz)));
  CUDA_TRY(cudaMalloc(&h->d_obs, n * h->n_obs * h->rsz));
  CUDA_TRY(cudaMalloc(&h->d_ref, n * (h->n_ref > 0 ? h->n_ref : 1) * h->rsz));
  CUDA_TRY(cudaMalloc(&h->d_rew, n * h->rsz));
  CUDA_TRY(cudaMalloc((void**)&h->d_term, n));
  CUDA_TRY(cudaMalloc((void**)&h->d_mask, n));
  return GEMB200_OK;
}

int gemb200_step_host(gemb200_handle* h, const void* action, void* obs_out, void* ref_out, void* reward_out, uint8_t* terminated_out) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  if (!action) return fail(GEMB200_E_INVALID, "action is NULL");
  DeviceGuard guard(h->cfg.device);
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  const size_t n = (size_t)h->cfg.n_envs;
  const size_t asz = (size_t)h->n_act * (h->cfg.finite ? sizeof(int32_t) : h->rsz);  // action bytes per env
  // Row-per-env buffers are contiguous per env range, so a large batch is cut into chunks that flow through three
  // streams: the D2H of chunk c overlaps the H2D + launch of chunk c+1 (PCIe is full duplex).  One API call = one RNG id.
  const bool pipelined = h->cfg.layout == GEMB200_LAYOUT_AOS && n >= (size_t)1 << 16;
  static const int chunks_env = [] { const char* e = std::getenv("GEMB200_HOST_CHUNKS"); return e ? std::atoi(e) : 0; }();  // experiment knob
  const int nchunk = pipelined ? (chunks_env > 0 ? chunks_env : 4) : 1;  // 2..16 chunks measure the same (PCIe D2H bound, ~49 GB/s); 4 keeps the copy count low
  const size_t per = pipelined ? ((n / nchunk + 255) / 256) * 256 : n;
  bool first = true;
  for (int c = 0; c < nchunk; ++c) {
    const size_t b = (size_t)c * per, e = (b + per < n) ? b + per : n;
    if (b >= e) break;
    cudaStream_t st = pipelined ? h->hpipe[c % 3] : h->hstream;
    CUDA_TRY(cudaMemcpyAsync((char*)h->d_act + b * asz, (const char*)action + b * asz, (e - b) * asz, cudaMemcpyHostToDevice, st));
    rc = do_step(h, h->d_act, obs_out ? h->d_obs : nullptr, (ref_out && h->n_ref) ? h->d_ref : nullptr, reward_out ? h->d_rew : nullptr,
                 terminated_out ? h->d_term : nullptr, st, (int)b, (int)e, first);
    first = false;
    if (rc) return rc;
    if (pipelined) {
      const size_t os = (size_t)h->n_obs * h->rsz, rs = (size_t)h->n_ref * h->rsz;
      if (obs_out) CUDA_TRY(cudaMemcpyAsync((char*)obs_out + b * os, (char*)h->d_obs + b * os, (e - b) * os, cudaMemcpyDeviceToHost, st));
      if (ref_out && h->n_ref) CUDA_TRY(cudaMemcpyAsync((char*)ref_out + b * rs, (char*)h->d_ref + b * rs, (e - b) * rs, cudaMemcpyDeviceToHost, st));
      if (reward_out) CUDA_TRY(cudaMemcpyAsync((char*)reward_out + b * h->rsz, (char*)h->d_rew + b * h->rsz, (e - b) * h->rsz, cudaMemcpyDeviceToHost, st));
      if (terminated_out) CUDA_TRY(cudaMemcpyAsync(terminated_out + b, h->d_term + b, e - b, cudaMemcpyDeviceToHost, st));
    }
  }
  if (pipelined) {
    for (int k = 0; k < 3; ++k) CUDA_TRY(cudaStreamSynchronize(h->hpipe[k]));
    return GEMB200_OK;
  }
  cudaStream_t st = h->hstream;
  if (obs_out) CUDA_TRY(cudaMemcpyAsync(obs_out, h->d_obs, n * h->n_obs * h->rsz, cudaMemcpyDeviceToHost, st));
  if (ref_out && h->n_ref) CUDA_TRY(cudaMemcpyAsync(ref_out, h->d_ref, n * h->n_ref * h->rsz, cudaMemcpyDeviceToHost, st));
  if (reward_out) CUDA_TRY(cudaMemcpyAsync(reward_out, h->d_rew, n * h->rsz, cudaMemcpyDeviceToHost, st));
  if (terminated_out) CUDA_TRY(cudaMemcpyAsync(terminated_out, h->d_term, n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return GEMB200_OK;
}

int gemb200_reset_host(gemb200_handle* h, const uint8_t* reset_mask, void* obs_out, void* ref_out) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  int rc = ensure_host_buffers(h);
  if (rc) return rc;
  const size_t n = (size_t)h->cfg.n_envs;
  cudaStream_t st = h->hstream;
  if (reset_mask) CUDA_TRY(cudaMemcpyAsync(h->d_mask, reset_mask, n, cudaMemcpyHostToDevice, st));
  if (reset_mask && (obs_out || ref_out)) {
    // unmasked envs keep the caller's previous values: pre-load the device staging buffers with them
    if (obs_out) CUDA_TRY(cudaMemcpyAsync(h->d_obs, obs_out, n * h->n_obs * h->rsz, cudaMemcpyHostToDevice, st));
    if (ref_out && h->n_ref) CUDA_TRY(cudaMemcpyAsync(h->d_ref, ref_out, n * h->n_ref * h->rsz, cudaMemcpyHostToDevice, st));
  }
  rc = do_reset(h, reset_mask ? h->d_mask : nullptr, obs_out ? h->d_obs : nullptr, (ref_out && h->n_ref) ? h->d_ref : nullptr, st);
  if (rc) return rc;
  if (obs_out) CUDA_TRY(cudaMemcpyAsync(obs_out, h->d_obs, n * h->n_obs * h->rsz, cudaMemcpyDeviceToHost, st));
  if (ref_out && h->n_ref) CUDA_TRY(cudaMemcpyAsync(ref_out, h->d_ref, n * h->n_ref * h->rsz, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return GEMB200_OK;
}

int gemb200_get_ode_state(gemb200_handle* h, double* ode_out, void* stream) {
  if (!h || !ode_out) return fail(GEMB200_E_INVALID, "NULL argument");
  DeviceGuard guard(h->cfg.device);
  const int n = h->cfg.n_envs, grid = (n + 255) / 256;
  if (h->cfg.dtype == GEMB200_F32) get_ode_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)h->d_st, (const float*)h->d_stc, h->d_eps, ode_out, n, h->nx, h->n_ref, h->has_eps);
  else get_ode_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)h->d_st, (const double*)h->d_stc, h->d_eps, ode_out, n, h->nx, h->n_ref, h->has_eps);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GEMB200_OK;
}
int gemb200_set_ode_state(gemb200_handle* h, const double* ode_in, void* stream) {
  if (!h || !ode_in) return fail(GEMB200_E_INVALID, "NULL argument");
  DeviceGuard guard(h->cfg.device);
  const int n = h->cfg.n_envs, grid = (n + 255) / 256;
  if (h->cfg.dtype == GEMB200_F32) set_ode_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)h->d_st, (float*)h->d_stc, h->d_eps, ode_in, n, h->nx, h->n_ref, h->has_eps);
  else set_ode_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)h->d_st, (double*)h->d_stc, h->d_eps, ode_in, n, h->nx, h->n_ref, h->has_eps);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GEMB200_OK;
}
int gemb200_get_reference(gemb200_handle* h, double* ref_out, void* stream) {
  if (!h || !ref_out) return fail(GEMB200_E_INVALID, "NULL argument");
  if (h->n_ref == 0) return GEMB200_OK;
  DeviceGuard guard(h->cfg.device);
  const int n = h->cfg.n_envs, grid = (n + 255) / 256;
  if (h->cfg.dtype == GEMB200_F32) get_ref_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)h->d_st, ref_out, n, h->nx, h->n_ref);
  else get_ref_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)h->d_st, ref_out, n, h->nx, h->n_ref);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GEMB200_OK;
}
int gemb200_set_reference(gemb200_handle* h, const double* ref_in, void* stream) {
  if (!h || !ref_in) return fail(GEMB200_E_INVALID, "NULL argument");
  if (h->n_ref == 0) return GEMB200_OK;
  DeviceGuard guard(h->cfg.device);
  const int n = h->cfg.n_envs, grid = (n + 255) / 256;
  if (h->cfg.dtype == GEMB200_F32) set_ref_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)h->d_st, ref_in, n, h->nx, h->n_ref);
  else set_ref_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)h->d_st, ref_in, n, h->nx, h->n_ref);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GEMB200_OK;
}

// checkpoint blob: [header][hot records][cold records][eps][sw][dead-time queue]...  The header pins the blob to the configuration that
// wrote it: a blob of equal size from another motor / seed / tau / generator set is refused instead of being reinterpreted.
struct Section { void* ptr; size_t bytes; };
static int sections(gemb200_handle* h, Section* s) {
  const size_t n = (size_t)h->cfg.n_envs;
  int k = 0;
  s[k++] = {h->d_st, n * (h->NH > 0 ? h->NH : 1) * h->rsz};
  s[k++] = {h->d_stc, n * h->NC * h->rsz};
  if (h->d_eps) s[k++] = {h->d_eps, n * sizeof(double)};
  if (h->d_sw) s[k++] = {h->d_sw, n * sizeof(uint16_t)};
  if (h->d_fifo) s[k++] = {h->d_fifo, n * h->cfg.dead_time_steps * h->fifo_dim * h->rsz};
  if (h->d_obsv) s[k++] = {h->d_obsv, n * 4 * h->rsz};
  if (h->d_sup) s[k++] = {h->d_sup, n * 2 * h->rsz};
  if (h->d_supph) s[k++] = {h->d_supph, n * sizeof(double)};
  if (h->d_swst) s[k++] = {h->d_swst, n * 2 * h->cfg.n_ref * sizeof(uint32_t)};
  if (h->d_kenv) s[k++] = {h->d_kenv, n * sizeof(uint32_t)};
  if (h->d_imprev) s[k++] = {h->d_imprev, n * 2 * h->rsz};
  return k;  // (the per-env parameter table is configuration, not state: re-apply gemb200_set_env_params after a load)
}
struct CheckpointHeader {
  char magic[8];          // "GEMB200C"
  int32_t abi, dtype, n_envs, n_sections, nh, nc, reserved[2];
  uint64_t config_hash, payload_bytes, gstep, n_steps;
};
// FNV-1a over the configuration, leaving out what may legitimately differ between writer and reader: the device ordinal and the
// host pointer of the speed-profile table (whose CONTENT is hashed at create)
static uint64_t config_hash(const gemb200_handle* h) {
  gemb200_config c = h->cfg;
  c.device = 0;
  c.ext_speed_table = nullptr;
  uint64_t x = 1469598103934665603ull;
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&c);
  for (size_t i = 0; i < sizeof(c); ++i) { x ^= b[i]; x *= 1099511628211ull; }
  x ^= h->ext_hash; x *= 1099511628211ull;
  return x;
}
static void make_header(gemb200_handle* h, CheckpointHeader* hd) {
  std::memset(hd, 0, sizeof(*hd));
  std::memcpy(hd->magic, "GEMB200C", 8);
  hd->abi = GEMB200_ABI_VERSION; hd->dtype = h->cfg.dtype; hd->n_envs = h->cfg.n_envs; hd->nh = h->NH; hd->nc = h->NC;
  Section s[16];
  hd->n_sections = sections(h, s);
  for (int i = 0; i < hd->n_sections; ++i) hd->payload_bytes += s[i].bytes;
  hd->config_hash = config_hash(h);
  hd->gstep = h->gstep; hd->n_steps = h->n_steps;
}
int64_t gemb200_checkpoint_size(gemb200_handle* h) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  CheckpointHeader hd;
  make_header(h, &hd);
  return (int64_t)sizeof(hd) + (int64_t)hd.payload_bytes;
}
int gemb200_checkpoint_save(gemb200_handle* h, void* host_blob) {
  if (!h || !host_blob) return fail(GEMB200_E_INVALID, "NULL argument");
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(cudaDeviceSynchronize());
  if (h->dev_clock) { int rc = pull_clock(h, nullptr); if (rc) return rc; }  // the header carries the clock
  char* b = (char*)host_blob;
  CheckpointHeader hd;
  make_header(h, &hd);
  std::memcpy(b, &hd, sizeof(hd)); b += sizeof(hd);
  Section s[16];
  const int k = sections(h, s);
  for (int i = 0; i < k; ++i) { CUDA_TRY(cudaMemcpy(b, s[i].ptr, s[i].bytes, cudaMemcpyDeviceToHost)); b += s[i].bytes; }
  return GEMB200_OK;
}
int gemb200_checkpoint_load(gemb200_handle* h, const void* host_blob) {
  if (!h || !host_blob) return fail(GEMB200_E_INVALID, "NULL argument");
  DeviceGuard guard(h->cfg.device);
  CheckpointHeader want, got;
  make_header(h, &want);
  std::memcpy(&got, host_blob, sizeof(got));
  if (std::memcmp(got.magic, want.magic, 8) != 0) return fail(GEMB200_E_INVALID, "checkpoint: not a gemb200 checkpoint blob (bad magic)");
  if (got.abi != want.abi) return fail(GEMB200_E_ABI, "checkpoint: written by another ABI version");
  if (got.dtype != want.dtype || got.n_envs != want.n_envs || got.n_sections != want.n_sections || got.nh != want.nh || got.nc != want.nc ||
      got.payload_bytes != want.payload_bytes)
    return fail(GEMB200_E_INVALID, "checkpoint: dtype / n_envs / record layout differ from this handle");
  if (got.config_hash != want.config_hash)
    return fail(GEMB200_E_INVALID, "checkpoint: written by a handle with a different configuration (motor, parameters, seed, tau, generators, ...)");
  CUDA_TRY(cudaDeviceSynchronize());
  const char* b = (const char*)host_blob + sizeof(got);
  h->gstep = got.gstep; h->n_steps = got.n_steps;
  Section s[16];
  const int k = sections(h, s);
  for (int i = 0; i < k; ++i) { CUDA_TRY(cudaMemcpy(s[i].ptr, b, s[i].bytes, cudaMemcpyHostToDevice)); b += s[i].bytes; }
  if (h->dev_clock) { int rc = push_clock(h, nullptr); if (rc) return rc; CUDA_TRY(cudaDeviceSynchronize()); }
  return GEMB200_OK;
}

// ElectricMotorEnvironment.reset(seed) -> _seed(seed) re-seeds every component (core.py:300-319, utils / RandomComponent.seed): a handle
// re-keyed with `seed` behaves exactly like a freshly created one with that seed — call ids, step clock, dead-time ring, switching
// states and every other persistent array start over — so equal seeds give identical episodes.
int gemb200_reseed(gemb200_handle* h, uint64_t seed, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  h->cfg.seed = seed;
  h->gstep = 0; h->n_steps = 0;
  if (h->dev_clock) { int rc = push_clock(h, st); if (rc) return rc; }
  for (int r = 0; r < 10; ++r) {
    const uint32_t lo = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u, hi = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    h->pf.rk[r][0] = lo; h->pf.rk[r][1] = hi; h->pd.rk[r][0] = lo; h->pd.rk[r][1] = hi;
  }
  h->pf.seed_lo = h->pd.seed_lo = (uint32_t)seed; h->pf.seed_hi = h->pd.seed_hi = (uint32_t)(seed >> 32);
  Section s[16];
  const int k = sections(h, s);
  for (int i = 0; i < k; ++i) {
    CUDA_TRY(cudaMemsetAsync(s[i].ptr, 0, s[i].bytes, st));
  }
  int rc = fill_imprev(h, st);
  if (rc) return rc;
  return do_reset(h, nullptr, nullptr, nullptr, st);
}

// Device-resident clock on / off (include/gemb200.h).  On: the host counters are uploaded; off: they are read back (synchronises `stream`).
int gemb200_set_device_clock(gemb200_handle* h, int32_t enable, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (enable && !h->dev_clock) {
    if (!h->d_clock) CUDA_TRY(cudaMalloc(&h->d_clock, 8 * sizeof(uint32_t)));
    int rc = push_clock(h, st);
    if (rc) return rc;
    h->dev_clock = true;
  } else if (!enable && h->dev_clock) {
    int rc = pull_clock(h, st);
    if (rc) return rc;
    h->dev_clock = false;
  }
  return GEMB200_OK;
}
int gemb200_get_clock(gemb200_handle* h, uint64_t* call_id, uint64_t* n_steps, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  if (h->dev_clock) { int rc = pull_clock(h, (cudaStream_t)stream); if (rc) return rc; }
  if (call_id) *call_id = h->gstep;
  if (n_steps) *n_steps = h->n_steps;
  return GEMB200_OK;
}

int64_t gemb200_launch_count(gemb200_handle* h) { return h ? h->launches : 0; }
int gemb200_kernel_time_begin(gemb200_handle* h, void* stream) {
  if (!h) return fail(GEMB200_E_INVALID, "handle is NULL");
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(cudaEventRecord(h->ev0, (cudaStream_t)stream));
  return GEMB200_OK;
}
int gemb200_kernel_time_end(gemb200_handle* h, void* stream, float* ms_out) {
  if (!h || !ms_out) return fail(GEMB200_E_INVALID, "NULL argument");
  DeviceGuard guard(h->cfg.device);
  CUDA_TRY(cudaEventRecord(h->ev1, (cudaStream_t)stream));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
  return GEMB200_OK;
}

}  // extern "C"
