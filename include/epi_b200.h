/* epi_b200.h — C-ABI of the B200-native batched epidemic step (libepi_b200.so).
 *
 * The reference (Nikunj-Gupta/Epipolicy_RL) has no FFI for this path: the seam is the Python
 * method call `Epidemic.get_next_state / Epidemic.step / Epidemic.reset`
 * (epipolicy/core/epidemic.py:82-116) driven by `EpiEnv.step / reset`
 * (epipolicy_environment.py:39-66). These entry points are what a ctypes binding placed at that
 * seam would call (INTEGRATION.md shows the binding). Plain pointers and sizes only.
 *
 * Conventions: every function returns 0 on success, a negative EPI_E_* code on failure and never
 * throws; `epi_last_error(h)` (or `epi_last_error(NULL)` after a failed epi_create) gives the text.
 * One handle = one CUDA device; calls on a handle are serialised by the caller; work is
 * stream-ordered on the `cuda_stream` argument (a `cudaStream_t` cast to void*, NULL = default).
 * "device" pointers are CUDA device memory owned by the caller (e.g. torch tensors).
 */
#ifndef EPI_B200_H
#define EPI_B200_H
#include <stddef.h>
#include <stdint.h>

#include "epi_desc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct epi_engine epi_t;

enum { EPI_OK = 0, EPI_E_INVALID = -1, EPI_E_UNSUPPORTED = -2, EPI_E_CUDA = -3, EPI_E_NOMEM = -4 };
enum { EPI_FP64 = 0, EPI_FP32 = 1 }; /* fp64 parity mode (1e-9) / fp32 throughput mode (1e-4) */

typedef struct EpiInfo {
  int32_t n_envs, precision, device;
  int32_t obs_dim;     /* C*L*G, observation = current_comp.flatten() in [C,L,G] order */
  int32_t action_dim;  /* A */
  int32_t ncp, n_itv;  /* NCP, I */
  int32_t horizon;
  int32_t n_flat;      /* length of the reference's flat ODE state */
  int32_t n_acc, n_slot, n_lc, n_groups_inf, n_cls;
  int32_t team_kind;   /* 0 = warp per env, 1 = CTA per env, 2 = cell kernel (lane per (env, locale, group)) */
  int32_t threads, grid, envs_per_cta;
  int32_t specialized; /* 1 once a scenario-specialised build of the cell kernel is attached */
  int32_t reserved;
  int64_t smem_bytes, small_bytes, big_bytes, scratch_bytes, state_bytes_per_env;
  int64_t mv_len;      /* floats of a move table per env (state selector 8): n_slot*L*G, or the number of general move
                          entries when the scenario has cross-locale moves */
} EpiInfo;

/* Epidemic.__init__ + make_setting (core/epidemic.py:36-80) for n_envs independent instances of one
 * scenario on CUDA device `device`. `desc` holds host pointers and is copied. */
int epi_create(const EpiDesc* desc, int n_envs, int device, int precision, epi_t** out);
void epi_destroy(epi_t* h);
const char* epi_last_error(const epi_t* h);
int epi_info(const epi_t* h, EpiInfo* out);

/* Epidemic.reset (core/epidemic.py:104-106) for the envs whose mask byte is non-zero (mask NULL = all).
 * obs_out: device [n_envs, obs_dim] of the handle's real type (double for EPI_FP64, float for EPI_FP32) or NULL. */
int epi_reset(epi_t* h, const uint8_t* mask_dev, void* obs_out_dev, void* cuda_stream);

/* EpiEnv.step (epipolicy_environment.py:39-62) for the batch: `actions_dev` float32 [n_envs, action_dim]
 * in [0,1]; n_days simulated days with the same action fused in one launch (the reference's PERIOD = 7);
 * stops early for an env that reaches the horizon. Outputs (each may be NULL):
 *   obs_out_dev    real   [n_envs, obs_dim]
 *   reward_out_dev double [n_envs]   sum over the days of -sum(delta cumulative_cost) (epidemic.py:115)
 *   done_out_dev   uint8  [n_envs]   day counter >= horizon (no auto-reset, like the reference) */
int epi_step(epi_t* h, const float* actions_dev, int n_days, void* obs_out_dev, double* reward_out_dev,
             uint8_t* done_out_dev, void* cuda_stream);

/* Epidemic.step (core/epidemic.py:108-116) with an explicit action list, as Epidemic.run uses it with
 * the scenario's schedule (:125-128): full control-parameter vectors `cpv_dev` float32 [n_envs, ncp] and
 * `active_dev` uint8 [n_envs, n_itv] (1 = the intervention has an Act today). `schedule_programs` = 1
 * resolves locale selectors as the schedule's day-0 Acts do, 0 as locale_regex '*'. */
int epi_step_plan(epi_t* h, const float* cpv_dev, const uint8_t* active_dev, int schedule_programs, int n_days,
                  void* obs_out_dev, double* reward_out_dev, uint8_t* done_out_dev, void* cuda_stream);

/* Per-day outputs of epi_run_plan, all device pointers, each may be NULL. Rows of days an env does not reach (horizon)
 * are left untouched. */
typedef struct EpiRunOut {
  void* obs_days;      /* real   [n_days, n_envs, obs_dim]      the day's end state, as epi_step's obs_out */
  double* reward_days; /* double [n_days, n_envs]               the day's reward (epidemic.py:115) */
  double* x_days;      /* double [n_days, n_envs, obs_dim]      the state at full precision (reporter input) */
  void* acc_days;      /* real   [n_days, n_envs, n_acc * L*G]  live cumulative components (epi_get_state_dev what = 10) */
  float* mult_days;    /* float  [n_days, n_envs, n_cls]        the day's class multipliers (what = 7) */
} EpiRunOut;

/* Epidemic.run (core/epidemic.py:118-134) for a batch of what-if schedules in ONE launch: day d of the launch takes
 * its Acts from row d of `cpv_days_dev` float32 [n_days, n_envs, ncp] and `active_days_dev` uint8 [n_days, n_envs, n_itv]
 * (NULL: every intervention has an Act every day) and emits the day's results into `days_out` (may be NULL). The final
 * outputs are those of epi_step_plan with the same n_days. Equivalent to n_days calls of epi_step_plan(.., n_days = 1):
 * bit-identical in fp64; the fp32 mode reuses the last stage of a day as the first of the next when the parameters
 * did not move, as epi_step does. */
int epi_run_plan(epi_t* h, const float* cpv_days_dev, const uint8_t* active_days_dev, int schedule_programs, int n_days,
                 const EpiRunOut* days_out, void* obs_out_dev, double* reward_out_dev, uint8_t* done_out_dev,
                 void* cuda_stream);

/* The same step through HOST buffers (what a CPU-side caller such as SB3's DummyVecEnv sees):
 * pinned staging, H2D of the actions, the fused launch, D2H of obs/reward/done, then a stream sync. */
int epi_step_host(epi_t* h, const float* actions_host, int n_days, void* obs_out_host, double* reward_out_host,
                  uint8_t* done_out_host);

/* Second destination for the results of every following epi_step / epi_step_plan (any pointer may be NULL; all NULL
 * unbinds): the kernel epilogue stores obs [n_envs, obs_dim] real, reward [n_envs] double and done [n_envs] uint8 there
 * as well. Meant for the learner rank's receive buffer mapped into this process (peer memory over NVLink): the stores
 * are the gather of SURVEY.md 8(e) (epipolicy_environment.py:56-62 returns obs/reward/done to the single caller; here
 * the caller that learns sits on another GPU), overlapped with the environments still being stepped. The caller
 * orders the reader behind the launch (stream + cross-rank barrier). */
int epi_bind_mirror(epi_t* h, void* obs_dev, double* reward_dev, uint8_t* done_dev);

/* The library's own pinned host buffers of the host path: actions float32 [n_envs, action_dim], obs real
 * [n_envs, obs_dim], reward double [n_envs], done uint8 [n_envs]. A caller that fills `*actions` and passes these
 * very pointers to epi_step_host gets the results in place, without the staging copies. Valid until epi_destroy;
 * overwritten by the next epi_step_host. */
int epi_host_buffers(epi_t* h, float** actions, void** obs, double** reward, uint8_t** done);

/* State access for checkpoint/resume and parity injection. `what`:
 *   0 X [n_envs, obs_dim] double        1 cumulative_cost [n_envs, I*L] double
 *   2 day counter [n_envs] int32        3 reference-layout flat observable [n_envs, n_flat] double (get only)
 *   4 trace [n_envs, 4] int32: nfev, accepted steps, rejected steps, status of the last simulated day
 *   5 last error norm [n_envs] double   6 yesterday's cost rate [n_envs, I*L] double
 *   7 yesterday's class multipliers [n_envs, n_cls] float    8 yesterday's move table [n_envs, mv_len] float
 *   9 (h, error_norm) of each attempted RK45 step of the last day [n_envs, 64, 2] double (get only; recording
 *     starts with the first query)
 * Host buffers; sizes are checked. */
int epi_get_state(epi_t* h, int what, void* out_host, size_t bytes);
int epi_set_state(epi_t* h, int what, const void* in_host, size_t bytes);

/* Device-side copies of the state (stream-ordered, no host round trip) for consumers that stay on the GPU - the
 * reporter quantities of core/reporter.py:45-204 for all envs (epipolicy_rl_b200/reporter_device.py). `what`:
 *   0 X double [n_envs, obs_dim]   1 cumulative_cost double [n_envs, I*L]   7 class multipliers float [n_envs, n_cls]
 *   10 live cumulative_move / gain / lost components, real [n_envs, n_acc, L*G] (component a sits at offset
 *      plan table 0 [a] * L*G of the reference's flat observable, obj/observable.py:24-25) */
int epi_get_state_dev(epi_t* h, int what, void* out_dev, size_t bytes, void* cuda_stream);
/* Static tables of the compiled scenario (host buffers, sizes checked): 0 acc_flat int32 [n_acc]; 1 locale class of
 * every locale int32 [L]; 2 default model parameters float [n_lc, P, F, G]; 3 their multiplier classes int32;
 * 4 default facility_timespent float [n_lc, F, G]; 5 their multiplier classes int32. With table 7 of the state
 * (class multipliers, class 0 = identity) they give a state's parameters: value * mult[class]
 * (obj/parameter.py:53-66). */
int epi_get_plan(const epi_t* h, int what, void* out_host, size_t bytes);

/* Scenario specialisation of the cell kernel (the reference JIT-compiles its RHS with numba; this is the
 * analogue). epi_spec_source writes a generated CUDA header (NUL-terminated, `*need` = bytes required)
 * holding the scenario's per-cell program as straight-line code; the caller compiles
 * csrc/epi_spec_main.cu with -DEPI_SPEC_HEADER="that file" into a shared object (spec.py does it with nvcc,
 * caching by the key inside the header) and hands the path to epi_spec_attach, which checks the key and
 * routes subsequent launches to it. Without an attached build the interpreted cell kernel runs. */
int epi_spec_source(epi_t* h, char* buf, size_t cap, size_t* need);
/* the same text from a descriptor alone: needs no CUDA device (ahead-of-time builds) */
int epi_spec_source_desc(const EpiDesc* desc, int precision, char* buf, size_t cap, size_t* need);
int epi_spec_attach(epi_t* h, const char* so_path);

/* number of kernel launches issued by this handle so far (bench bookkeeping) */
int64_t epi_launch_count(const epi_t* h);

#ifdef __cplusplus
}
#endif
#endif /* EPI_B200_H */
