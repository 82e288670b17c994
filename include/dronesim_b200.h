/*
 * dronesim_b200 - C ABI of the B200-native batched dynamics + INDI core.
 *
 * The reference (enac-drones/dronesim) is pure Python and has no FFI of its own; its only seam
 * is the Python class API.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes stubs a
 * reference maintainer would add to route BaseAviary.step / INDIControl.computeControl here.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every pointer is either HOST or DEVICE as documented.
 *   - every function returns 0 (DS_OK) or a non-zero ds_status; nothing throws.
 *   - one handle per CUDA device; calls on a handle are stream-ordered, asynchronous, and
 *     not thread-safe.  After one warm-up call (lazy allocations) ds_step / ds_physics_step / ds_control_* only enqueue
 *     kernels, so they can be captured into a CUDA graph and replayed; host-side bookkeeping (step_counter, the
 *     max_steps predicate, the noise substep index, the trajectory log) advances at capture time only.  The handle owns all device state; ds_views() lends pointers that stay
 *     valid until ds_destroy().
 *   - vehicles are numbered v = env * drones_per_env + slot (env-major).  Quaternions are
 *     xyzw (PyBullet / dronesim/utils/math.py:6,25).  Commands are PWM in [MIN_PWM, MAX_PWM].
 *   - there is NO CPU fallback: without a CUDA device ds_create() fails with DS_ERR_CUDA.
 */
#ifndef DRONESIM_B200_H
#define DRONESIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_ABI_VERSION 2
#define DS_MAX_ROTORS 6
#define DS_MAX_TYPES 8
#define DS_MAX_DRONES_PER_ENV 32
#define DS_OBS_STRIDE 22 /* 16 + DS_MAX_ROTORS floats per vehicle (BaseAviary.py:780-790) */
#define DS_NUM_STATS 16

typedef struct ds_handle ds_handle;

typedef enum ds_status {
  DS_OK = 0,
  DS_ERR_INVALID = 1,     /* bad argument */
  DS_ERR_CUDA = 2,        /* CUDA runtime error (see ds_last_cuda_error) */
  DS_ERR_STATE = 3,       /* call order (e.g. step before set_types / reset) */
  DS_ERR_UNSUPPORTED = 4  /* valid request outside what this build implements */
} ds_status;

/* integrators: DS_INTEG_RPY is the literal BaseAviary._dynamics update (BaseAviary.py:1809-1817),
 * DS_INTEG_QUAT is rigid-body Newton-Euler about the composite CoM with a quaternion update. */
enum { DS_INTEG_QUAT = 0, DS_INTEG_RPY = 1 };

/* aerodynamic add-ons = the Physics enum of the reference (BaseAviary.py:41-49):
 * pyb_gnd -> GROUND, pyb_drag -> DRAG, pyb_dw -> DOWNWASH, pyb_gnd_drag_dw -> all three. */
enum { DS_FLAG_GROUND = 1u, DS_FLAG_DRAG = 2u, DS_FLAG_DOWNWASH = 4u, DS_FLAG_STATS = 8u,
       /* diagnostics: evaluate every ORDERED downwash pair even where the symmetric kernel variant applies
        * (16 drones per env, all types sharing dw_coeff_2 / dw_coeff_3); results agree to FP32 summation order */
       DS_FLAG_DW_ORDERED_PAIRS = 16u,
       /* diagnostics: a single-type swarm runs the mixed-swarm kernel variant (per-type tables in shared memory) instead
        * of the homogeneous one (its type table in the kernel parameters); results are bit-identical */
       DS_FLAG_TYPES_IN_SMEM = 32u,
       /* the ground plane of the reference's PyBullet world as a hard floor at ds_config.ground_plane_z (off by default:
        * the reference's explicit-dynamics formulas have no contact) */
       DS_FLAG_GROUND_PLANE = 64u,
       /* diagnostics: every device buffer of the handle is allocated between two 4 KiB guard bands of a known pattern;
        * ds_debug_check_redzones counts the guard bytes a kernel has overwritten (the library's own stray-store check) */
       DS_FLAG_DEBUG_REDZONES = 128u };

/* control laws: which reference controller class flies the type */
enum { DS_LAW_QUAD = 0 /* INDIControl.py */, DS_LAW_6DOF = 1 /* INDIControl_6DOF.py */ };

/* done bits (per vehicle, sticky).  GOAL follows examples/fly_INDI_TrajectoryTrack.py:249-250;
 * FLOOR and TIME are extensions, off unless configured. */
enum { DS_DONE_GOAL = 1u, DS_DONE_FLOOR = 2u, DS_DONE_TIME = 4u };

/* control / physics order inside ds_step */
enum {
  DS_ORDER_PHYSICS_THEN_CONTROL = 0, /* examples/fly_INDI.py:217-245 (CtrlAviary loop) */
  DS_ORDER_CONTROL_THEN_PHYSICS = 1  /* VelocityAviary._preprocessAction (VelocityAviary.py:221-264) */
};

typedef struct ds_config {
  int32_t n_envs;          /* environments held by THIS handle (one GPU's shard) */
  int32_t drones_per_env;  /* NUM_DRONES of each env, 1..DS_MAX_DRONES_PER_ENV (BaseAviary.py:189) */
  int32_t substeps;        /* AGGR_PHY_STEPS (BaseAviary.py:187) */
  int32_t integrator;      /* DS_INTEG_* */
  uint32_t flags;          /* DS_FLAG_* */
  int32_t device;          /* CUDA device ordinal */
  float sim_freq;          /* SIM_FREQ, 240 (BaseAviary.py:185) */
  float gravity;           /* G = 9.8 (BaseAviary.py:182) */
  float neighbourhood_radius; /* NEIGHBOURHOOD_RADIUS (BaseAviary.py:190); INFINITY allowed */
  /* batched done predicate; all off reproduces the reference's constant False (CtrlAviary.py:282-293) */
  int32_t done_goal_enable;
  float goal[3];
  float goal_radius;       /* 0.3 in fly_INDI_TrajectoryTrack.py:249 */
  int32_t done_floor_enable;
  float z_min;
  int32_t max_steps;       /* DS_DONE_TIME when step_counter >= max_steps; 0 = off */
  int32_t env_offset;      /* global index of this shard's first env (multi-GPU bookkeeping only) */
  /* Extensions the north_star asks for beyond the reference; 0 = off = the reference's behaviour.  Quaternion
   * integrator only (DS_ERR_UNSUPPORTED with DS_INTEG_RPY). */
  float motor_tau;         /* s: first-order motor model, rpm += (1 - exp(-dt / tau)) (rpm_cmd - rpm) every substep;
                              0 = the static PWM -> RPM map (BaseAviary.py:1487-1490) */
  float acc_filter_hz;     /* Hz: first-order low-pass on the INDI angular-acceleration estimate; 0 = the raw finite
                              difference (the reference's filter is a commented placeholder, INDIControl.py:432-439) */
  int32_t reward_mode;     /* 0: constant -1 (CtrlAviary.py:267-278); 1: minus the env's mean |pos_e| of the last control step */
  /* Rotor noise of the live force models (BaseAviary.py:1429-1432, 1518-1525: N(0, 0.01) on every rotor thrust,
   * N(0, 0.001) on every reaction torque, plus the quad model's lateral / base terms :1528-1543), per substep.  The
   * reference draws it from the unseeded global numpy generator; here it is a counter-based stream (Philox-4x32-10,
   * Box-Muller) keyed by `noise_seed` and indexed by (global vehicle id, substep), so a run is reproducible and does
   * not depend on how the envs are sharded.  sigma = 0: off (all parity runs). */
  float noise_force_sigma;
  float noise_torque_sigma;
  uint64_t noise_seed;
  /* DS_FLAG_GROUND_PLANE: the plane the reference loads under the aviary (plane.urdf, BaseAviary.py:679-680), as an
   * inelastic frictionless stop: after every substep a centre of mass below ground_plane_z is put back on it and loses its
   * downward velocity.  A stand-in for Bullet's contact solver, enough for the take-off phases of the example scripts
   * (fly_INDI.py starts with all-zero controller commands and touches down before it lifts off). */
  float ground_plane_z;
  int32_t reserved0;
} ds_config;

/* Per-type constants.  Field sources: BaseAviary._parseURDFParameters (BaseAviary.py:2041-2140),
 * INDIControl._parseURDFControlParameters (INDIControl.py:55-106) and the URDF kinematic tree. */
typedef struct ds_type_params {
  int32_t n_u;   /* INDI_ACTUATOR_NR */
  int32_t n_v;   /* INDI_OUTPUT_NR */
  int32_t law;   /* DS_LAW_* */
  int32_t rotor_model; /* 0: quad, forces on links 0..n_u-1 + one base torque (BaseAviary.py:1477-1543); 1: morphing hexa
                          (:1389-1457); 2: quad whose TYPE carries "advanced": oblique-flow propeller model (:1493-1512,
                          1570-1644; quaternion integrator only) */
  double mass;          /* M (literal) or whole-tree mass */
  double J[9];          /* row-major inertia about the centre of mass, body axes */
  double r_com[3];      /* centre of mass in the base frame */
  double kf, km;        /* KF, KM */
  double rotor_pos[DS_MAX_ROTORS][3];   /* base frame */
  double rotor_axis[DS_MAX_ROTORS][3];  /* thrust direction */
  double torque_axis[DS_MAX_ROTORS][3]; /* reaction-torque direction */
  double rotor_spin[DS_MAX_ROTORS];     /* sign of KM*rpm^2 */
  double pwm2rpm_scale[DS_MAX_ROTORS], pwm2rpm_const[DS_MAX_ROTORS];
  double min_pwm[DS_MAX_ROTORS], max_pwm[DS_MAX_ROTORS];
  double gnd_eff_coeff, prop_radius, gnd_eff_h_clip;
  double drag_coeff[3];
  double dw_coeff[3];
  double kp_pos, kd_pos;                /* indi_guidance_gains */
  double att_gain[3], rate_gain[3];     /* indi_att_gains att / rate p,q,r */
  double G1[DS_MAX_ROTORS][DS_MAX_ROTORS]; /* [n_v][n_u] control effectiveness (URDF indi_1..n_v) */
  double alloc[DS_MAX_ROTORS][DS_MAX_ROTORS]; /* [n_u][n_v]: pinv(G1/0.05) (law QUAD, INDIControl.py:459)
                                                 or the first-iteration WLS matrix (law 6DOF, wls_alloc.py:190-259) */
  double wls_wv[DS_MAX_ROTORS];         /* Wv of INDIControl_6DOF.py:618 */
  double wls_gamma;                     /* gamma_sq of wls_alloc.py:125 */
  double init_cmd;                      /* controller reset: 0.0 (INDIControl.py:129) / 0.5 (INDIControl_6DOF.py:234) */
  double init_thrust;                   /* 0.0 (INDIControl.py:127) / 0.3 (INDIControl_6DOF.py:232) */
  double max_speed_kmh;                 /* MAX_SPEED_KMH (BaseAviary.py:2079): SPEED_LIMIT of VelocityAviary.py:92-94 */
  double adv_coeff[14];                 /* rotor_model 2: Data_section5_ObliqueFlow["mamr-8x4.5"] (propeller_database.py:537-552) */
  double adv_radius;                    /* propeller radius in metres (utils/utils.py:172-174) */
} ds_type_params;

/* Where the controller's set-points come from.  Every array is 16-byte aligned (rows are float4; DS_ERR_INVALID otherwise). */
typedef struct ds_targets {
  int32_t mode;          /* 0: per-vehicle device arrays; 1: shared waypoint table + per-vehicle counter;
                            2: velocity command (VelocityAviary._preprocessAction, VelocityAviary.py:221-264): `vel`
                               holds the raw action (x, y, z direction, fraction of the speed limit); the target
                               position / yaw are the vehicle's own, target_vel = SPEED_LIMIT |a3| unit(a012);
                            3: body-rate + thrust command (RPYTAviary._preprocessAction, RPYTAviary.py:180-193 ->
                               INDIControl._INDIRateControl, INDIControl.py:413-490): `vel` holds (p, q, r set-point, thrust);
                               quad-law types only (DS_ERR_UNSUPPORTED otherwise) */
  int32_t num_wp;        /* mode 1: rows in table */
  int32_t advance_wp;    /* mode 1: wp = wp+1 if wp < num_wp-1 else 0 after each control step (fly_INDI.py:242-245) */
  int32_t reserved;
  const float* pos_yaw;  /* mode 0: DEVICE [N][4] = target x,y,z,yaw */
  const float* vel;      /* mode 0: DEVICE [N][4] (xyz used) or NULL = zeros; mode 2: DEVICE [N][4] velocity action;
                            mode 3: DEVICE [N][4] rate / thrust action */
  const float* acc;      /* mode 0: DEVICE [N][4] (xyz used) or NULL = zeros */
  const float* table;    /* mode 1: DEVICE [num_wp][12] = pos xyz,yaw | vel xyz,0 | acc xyz,0 */
  const float* offset;   /* mode 1: DEVICE [N][4] additive position offset or NULL */
  const int32_t* wp;     /* mode 1: DEVICE [N] waypoint index of every vehicle for THIS control step, supplied by the
                            caller the way the reference scripts index TARGET_POS[wp_counters[j]] (fly_INDI.py:230-245),
                            or NULL = the resident per-vehicle counter (which advance_wp then advances).  Indices are
                            clamped to [0, num_wp). */
} ds_targets;

/* Borrowed device pointers of the resident state (structure of float4 arrays, each [n_pad]). */
typedef struct ds_state_views {
  int64_t n;             /* vehicles */
  int64_t n_pad;         /* allocated vehicles (n rounded up to the tile size) */
  float* pos_thrust;     /* [n_pad][4] pos x,y,z | controller last_thrust */
  float* quat;           /* [n_pad][4] x,y,z,w */
  float* vel_rpm;        /* [n_pad][4] vel x,y,z | sum of rpm of the last applied action (drag, BaseAviary.py:532) */
  float* omega_wp;       /* [n_pad][4] body rates p,q,r | waypoint counter (int32 bits) */
  float* lastvel_done;   /* [n_pad][4] controller last_vel xyz | done bits (uint32 bits) */
  float* lastrates_err;  /* [n_pad][4] controller last_rates pqr | |pos_e| of the last control step */
  float* cmd0123;        /* [n_pad][4] controller cmd / action, rotors 0..3 */
  float* cmd45;          /* [n_pad][2] rotors 4,5 */
  uint8_t* slot_type;    /* [drones_per_env] type id of each slot */
  int64_t step_counter;  /* BaseAviary.step_counter (BaseAviary.py:554) */
  float* rpm0123;        /* [n_pad][4] actual rotor speeds 0..3 (motor model on) or NULL */
  float* rpm45;          /* [n_pad][2] or NULL */
  float* ang_acc_filt;   /* [n_pad][4] filtered angular acceleration x,y,z (filter or motor model on) or NULL */
} ds_state_views;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* BaseAviary.__init__ (BaseAviary.py:129-198) for n_envs copies of the aviary. */
int ds_create(const ds_config* cfg, ds_handle** out);
void ds_destroy(ds_handle* h);
/* drone_model list -> self.drones (BaseAviary.py:219): HOST table + HOST slot->type map. */
int ds_set_types(ds_handle* h, const ds_type_params* types, int32_t n_types, const uint8_t* slot_type);
/* BaseAviary.reset/_housekeeping (BaseAviary.py:406-424, 640-714) + INDIControl.reset
 * (INDIControl.py:109-146).  HOST arrays: pos0 [N][3]; rpy0 [N][3] or NULL; vel0 [N][3] or NULL;
 * action0 [N][DS_MAX_ROTORS] = the action applied until the first control step (fly_INDI.py:214)
 * or NULL = zeros; wp0 [N] int32 or NULL = zeros. */
int ds_reset(ds_handle* h, const float* pos0, const float* rpy0, const float* vel0, const float* action0,
             const int32_t* wp0, void* stream);

/* BaseAviary.reset (BaseAviary.py:406-424) for SOME environments of the batch, on the device and without a host
 * synchronisation: every env e with mask_env[e] != 0 gets exactly the state ds_reset would give it (kinematics,
 * controller memory, done bits cleared, waypoint counter, and - when action0 is given - action0 as the first applied
 * action, fly_INDI.py:214); the other envs are not touched.  All pointers DEVICE: mask_env [n_envs] uint8; pos0 [N][3];
 * rpy0 / vel0 [N][3], action0 [N][DS_MAX_ROTORS], wp0 [N] or NULL; rows of unmasked envs are not read.  The time limit
 * (max_steps) of a reset env counts from this call.  Stream-ordered. */
int ds_reset_envs(ds_handle* h, const uint8_t* mask_env, const float* pos0, const float* rpy0, const float* vel0,
                  const float* action0, const int32_t* wp0, void* stream);

/* Checkpoint / resume: the resident state is the arrays ds_views exposes (copy them out and back in with any device
 * copy) plus the host-side step counter, which this call restores (BaseAviary.step_counter, BaseAviary.py:554; it also keys
 * the noise stream and the time limit).  Also clears the "first action pending" flag of a fresh ds_reset. */
int ds_set_step_counter(ds_handle* h, int64_t step_counter);

/* ---- the hot path ---------------------------------------------------------------------- */
/* n_control_steps x { K physics substeps with the held command ; one INDI evaluation } fused in
 * one kernel per control step: examples/fly_INDI.py:217-245 for all vehicles at once.  `order` picks the kernel
 * variant (DS_ORDER_*).  When a 6-DOF type is present every step kernel is followed by a tiny fix-up kernel that
 * solves the WLS allocation of the (rare) vehicles whose closed-form first iterate was infeasible
 * (wls_alloc.py:264) - the command they apply in the NEXT physics step; with DS_ORDER_CONTROL_THEN_PHYSICS and a
 * 6-DOF type the step runs as a control kernel followed by a physics kernel instead (same results). */
int ds_step(ds_handle* h, const ds_targets* tgt, int32_t n_control_steps, int32_t order, void* stream);
/* BaseAviary.step with an external action (BaseAviary.py:428-555): clip (CtrlAviary.py:258-263),
 * K substeps.  action: DEVICE [N][DS_MAX_ROTORS] PWM. */
int ds_physics_step(ds_handle* h, const float* action, void* stream);
/* INDIControl.computeControl on the resident state (INDIControl.py:154-227 /
 * INDIControl_6DOF.py:259-336).  Outputs (DEVICE, any may be NULL): cmd_out [N][DS_MAX_ROTORS],
 * pos_e_out [N][3], yaw_err_out [N].  The resident command is updated as well. */
int ds_control_step(ds_handle* h, const ds_targets* tgt, float control_timestep, float* cmd_out,
                    float* pos_e_out, float* yaw_err_out, void* stream);
/* BaseControl.computeControlFromState (BaseControl.py:61-103): same, but the kinematic state is
 * read from caller-provided aviary state vectors, DEVICE [N][DS_OBS_STRIDE]. */
int ds_control_from_state(ds_handle* h, const float* state, const ds_targets* tgt, float control_timestep,
                          float* cmd_out, float* pos_e_out, float* yaw_err_out, void* stream);
/* INDIControl._INDIRateControl (INDIControl.py:413-490), the RPYTAviary entry (RPYTAviary.py:180-193):
 * rate_thrust DEVICE [N][4] = p,q,r set-point, thrust. */
int ds_rate_control_step(ds_handle* h, const float* rate_thrust, float control_timestep, float* cmd_out,
                         void* stream);

/* ---- observation / bookkeeping --------------------------------------------------------- */
/* CtrlAviary._computeObs (CtrlAviary.py:212-232): obs DEVICE [N][DS_OBS_STRIDE] =
 * pos3 quat4 rpy3 vel3 ang_v_world3 last_clipped_action[6] (BaseAviary.py:780-790, zero padded);
 * neighbors DEVICE [N] uint32 = row of _getAdjacencyMatrix (BaseAviary.py:901-921) as a bitmask;
 * done_env DEVICE [n_envs] uint8; reward_env DEVICE [n_envs] float (constant -1, CtrlAviary.py:267-278).
 * Any pointer may be NULL. */
int ds_get_obs(ds_handle* h, float* obs, uint32_t* neighbors, uint8_t* done_env, float* reward_env, void* stream);
/* Arms per-env outputs of every following ds_step: done_env DEVICE [n_envs] uint8 and / or reward_env DEVICE [n_envs]
 * float (NULL, NULL disarms), written with the values ds_get_obs would return after that step.  When every env sits
 * inside one warp (drones_per_env divides 32) they are reduced with warp shuffles inside the fused kernel - no extra
 * launch; otherwise ds_step launches the observation kernel for them. */
int ds_set_env_outputs(ds_handle* h, uint8_t* done_env, float* reward_env);
int ds_views(ds_handle* h, ds_state_views* out);
/* Rollout statistics accumulated when DS_FLAG_STATS is set; synchronises the stream.
 * out[0]=control evaluations, [1]=sum |pos_e|^2, [2]=saturated rotor commands, [3]=WLS slow-path entries,
 * [4]=WLS non-convergences, [5]=non-finite states, [6]=min altitude, [7]=done vehicles. */
int ds_stats(ds_handle* h, double* host_out, int32_t n, void* stream);
int ds_stats_reset(ds_handle* h, void* stream);

/* ---- trajectory capture: dronesim/utils/Logger.py on the device --------------------------- */
/* Logger(logging_freq_hz, num_drones, duration_sec) + Logger.log (Logger.py:22-139): attach `n_vehicles` vehicles
 * (HOST ids, v = env * drones_per_env + slot); from then on every control step of ds_step and every ds_physics_step
 * appends one sample - the vehicle's aviary state vector (BaseAviary.py:780-790, what the examples pass to
 * logger.log) - to a DEVICE array laid out like Logger.states: [n_vehicles][DS_OBS_STRIDE][capacity].  Samples past
 * `capacity` are dropped; ds_reset rewinds the log; n_vehicles = 0 detaches. */
int ds_log_attach(ds_handle* h, const int32_t* vehicles, int32_t n_vehicles, int32_t capacity);
/* Copies the whole array to HOST host_states [n_vehicles][DS_OBS_STRIDE][capacity] (may be NULL) and the sample
 * times step_counter / SIM_FREQ to HOST host_timestamps [capacity] (may be NULL); *count_out = samples taken. */
int ds_log_read(ds_handle* h, float* host_states, double* host_timestamps, int32_t* count_out, void* stream);

/* ---- end-to-end convenience with HOST buffers ------------------------------------------ */
/* One control step driven from the host: copies HOST targets [N][4] (x,y,z,yaw) to the device,
 * runs ds_step, materialises obs and copies obs [N][DS_OBS_STRIDE] + done_env [n_envs] back to
 * HOST memory (pinned buffers give the best overlap).  Synchronises the stream before returning. */
int ds_step_host(ds_handle* h, const float* host_pos_yaw, float* host_obs, uint8_t* host_done_env, void* stream);

/* n_steps control steps driven from the host with the copies PIPELINED against the compute: the targets of
 * step i+1 (HOST [n_steps][N][4], pinned) travel to the device on a copy stream while step i runs, and the
 * per-env done flags of step i (HOST [n_steps][n_envs], pinned, or NULL) travel back while step i+1 runs.
 * Use when the set-points of the next steps are known ahead (the reference examples' TARGET_POS tables,
 * fly_INDI_TrajectoryTrack.py:133-160).  Every step still pays its own H2D / D2H; they just overlap.
 * Synchronises before returning. */
int ds_rollout_host(ds_handle* h, const float* host_pos_yaw, int32_t n_steps, uint8_t* host_done_env, void* stream);
/* The same pipeline fed the way the reference scripts feed their controllers: a waypoint table resident on the device
 * (tgt: mode 1, DEVICE table and optional DEVICE per-vehicle offsets) and, per control step, ONE index per vehicle
 * (host_wp HOST [n_steps][N] int32, pinned) - fly_INDI.py:230-245 passes TARGET_POS[wp_counters[j]].  4 bytes per vehicle
 * and step cross the bus instead of 16. */
int ds_rollout_host_table(ds_handle* h, const ds_targets* tgt, const int32_t* host_wp, int32_t n_steps,
                          uint8_t* host_done_env, void* stream);

/* ---- diagnostics ----------------------------------------------------------------------- */
/* The 6-DOF allocation alone: wls_alloc(v, MIN-cmd, MAX-cmd, G1/0.05, None, None, Wv, 1, None)
 * (INDIControl_6DOF.py:607-628 -> wls_alloc.py:125-350) for n independent problems of type type_id.
 * DEVICE v [n][6], cmd [n][6] -> du_out [n][6], iter_out [n] (iterations; negative = non-convergence),
 * w_out [n][6] or NULL (the final working set W in {-1, 0, +1}, wls_alloc.py:171, 284, 335-338).
 * force_slow != 0 skips the closed-form first iteration and always runs the FP64 active-set loop. */
int ds_debug_wls(ds_handle* h, int32_t type_id, const float* v, const float* cmd, float* du_out, int32_t* iter_out,
                 int32_t* w_out, int32_t n, int32_t force_slow, void* stream);

/* Handles created with DS_FLAG_DEBUG_REDZONES: synchronises the device and writes to *corrupted_bytes (HOST) how many
 * guard-band bytes around the handle's device buffers no longer hold the fill pattern (0 = no kernel stored outside
 * its buffers).  DS_ERR_UNSUPPORTED for handles created without the flag.  The reference has no counterpart (NumPy
 * raises IndexError on a stray index, BaseAviary.py:718-732 slices); this replaces it for the CUDA core. */
int ds_debug_check_redzones(ds_handle* h, int64_t* corrupted_bytes);

/* FP32 issue-rate micro-benchmark (bench bookkeeping: the FP32 roofline denominator, which
 * MEASURED_PEAKS.json does not carry).  Runs 8 independent FFMA chains per thread on every SM of
 * `device` and writes the best-of-5 rate in TFLOP/s (FMA = 2 FLOP) to *tflops_out (HOST). */
int ds_debug_fp32_peak(int32_t device, double* tflops_out);

/* ---- misc ------------------------------------------------------------------------------ */
const char* ds_strerror(int status);
int ds_last_cuda_error(ds_handle* h);
int ds_abi_version(void);
/* number of kernels this handle has launched since creation (bench bookkeeping) */
int64_t ds_launch_count(ds_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DRONESIM_B200_H */
