// Kernels of the dronesim_b200 core.
//
//  ds_step_kernel<INTEG, DW, NU6, WARPSYNC, MODE, FX, EXT>
//      MODE 0 (fused): K physics substeps, then one INDI evaluation per vehicle  (examples/fly_INDI.py:217-245)
//      MODE 1 (physics only): BaseAviary.step with an external action            (BaseAviary.py:428-555)
//      MODE 2 (fused): one INDI evaluation, then K substeps with the new command (VelocityAviary / RPYTAviary.step)
//      The order is a template parameter on purpose: with a run-time switch the control law is inlined at both call
//      sites (3,650 instead of 2,600 instructions) and the hot kernel runs 1.7 % (mixed swarm) to 9 % (hexa) slower.
//  ds_control_kernel<NU6>      INDIControl.computeControl on resident or external state
//  ds_obs_kernel               CtrlAviary._computeObs: state vector + adjacency bitmask + done / reward
//  ds_reset_kernel             BaseAviary._housekeeping + INDIControl.reset
//
// Grid: persistent CTAs looping over tiles of tile_v vehicles (whole environments per tile, so the
// downwash neighbour exchange never leaves the CTA); grid size = min(tiles, SMs x resident CTAs).
//
// Thread -> vehicle map: vehicles are env-major in memory (v = env * D + slot) and thread t of a tile
// handles local vehicle t, so a warp reads 512 contiguous bytes of every state array and (for D | 32)
// holds whole envs: the downwash snapshot is exchanged with __syncwarp only.  Sorting the threads of a
// tile by airframe class (warp-uniform control law / rotor count) was measured and rejected: it removed
// 2.6 % of the issued instructions (divergence only idles 4.5 % of the lanes in the 8 quad + 8 hexa env)
// but needed a CTA-wide barrier per substep and ran 7 % slower (profiles/r01_notes.md).
#pragma once
#include "ds_control.cuh"
#include "ds_device.cuh"
#include "ds_physics.cuh"

__device__ __forceinline__ void ds_load_types(const DsArgs& a, DsTypeDev* sh_types) {
  const int words = a.n_types * (int)(sizeof(DsTypeDev) / 16);
  const float4* src = reinterpret_cast<const float4*>(a.types);
  float4* dst = reinterpret_cast<float4*>(sh_types);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ float ds_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ds_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Rollout statistics: per-thread accumulators live in SHARED memory (one column per thread) so they cost no
// registers across the persistent tile loop; one warp-shuffle reduction + atomics per CTA at the end.
enum { ST_NCTRL = 0, ST_ERR2, ST_SAT, ST_WLS_SLOW, ST_WLS_FAIL, ST_NONFINITE, ST_MINZ, ST_DONE, ST_COUNT };

// One reduction per CTA (warp shuffles, then the four warp leaders through shared memory), then fire-and-forget
// reductions (RED, no return value to wait for) on the global accumulators: nothing at the end of a CTA waits for
// global memory.  Min altitude without a compare-and-swap loop: non-negative doubles order like their bit patterns
// taken as signed integers, negative ones in reverse, and every negative pattern is below every non-negative one as a
// signed integer but above it as an unsigned one - so a non-negative candidate is a signed min, a negative one an
// unsigned max, and either is correct against whatever the slot holds.
__device__ __forceinline__ void ds_flush_stats(float* sh_stat, double* stats) {
  const int t = threadIdx.x;
  float v[ST_COUNT];
#pragma unroll
  for (int i = 0; i < ST_COUNT; ++i) v[i] = sh_stat[i * DS_TILE + t];
#pragma unroll
  for (int i = 0; i < ST_COUNT; ++i) v[i] = (i == ST_MINZ) ? ds_warp_min(v[i]) : ds_warp_sum(v[i]);
  __syncthreads();  // every thread has read its column
  if ((t & 31) == 0) {
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) sh_stat[i * (DS_TILE / 32) + (t >> 5)] = v[i];
  }
  __syncthreads();
  if (t < ST_COUNT) {
    float r = sh_stat[t * (DS_TILE / 32)];
#pragma unroll
    for (int w = 1; w < DS_TILE / 32; ++w) {
      const float x = sh_stat[t * (DS_TILE / 32) + w];
      r = (t == ST_MINZ) ? fminf(r, x) : r + x;
    }
    if (t != ST_MINZ) {
      if (r != 0.f) atomicAdd(stats + t, (double)r);
    } else {
      const double d = (double)r + 0.0;  // -0.0 -> +0.0
      if (d >= 0.0) atomicMin(reinterpret_cast<long long*>(stats + t), __double_as_longlong(d));
      else atomicMax(reinterpret_cast<unsigned long long*>(stats + t), (unsigned long long)__double_as_longlong(d));
    }
  }
}

// Out-of-line copy for the kernels with a long substep loop: inlined, the epilogue changes the register allocation of
// the whole kernel (measured on the 16-drone mixed swarm: +0.8 % per step); the short single-vehicle-env kernels keep it inline
// (the call costs them 0.7 %).
static __device__ __noinline__ void ds_flush_stats_call(float* sh_stat, double* stats) { ds_flush_stats(sh_stat, stats); }

// ---------------------------------------------------------------------------------------------
// Tile staging: bulk asynchronous copies (TMA unit, cp.async.bulk) global -> shared, two stages per CTA.
//
// One thread per CTA issues one bulk copy per state array for the tile the CTA will process TWO iterations
// later (the slabs are contiguous: vehicles are env-major and a tile is a run of whole envs); completion is
// tracked by one mbarrier per stage (expect_tx = bytes of the tile).  All threads wait on the stage's mbarrier
// at the top of a tile and read their vehicle's rows with conflict-free LDS.128, so the DRAM latency of tile
// i + 1 is hidden behind the K substeps of tile i instead of being exposed at the tile start (physics state)
// and again at the control phase (controller memory, targets).  The optional target velocity / acceleration
// arrays are not staged (shared-memory budget); they are pulled into L2 with cp.async.bulk.prefetch.L2.
//
// Stage layout (T = DS_TILE rows of 16 B each unless noted):
//   POS | QUAT | VEL | OM | LV | LR | C0 | TG (per-vehicle target: pos+yaw, velocity action, rate/thrust or offset) | C1 (8 B rows)
// MODE 1 (physics only) stages POS..LV.
// ---------------------------------------------------------------------------------------------
enum { SG_POS = 0, SG_QUAT = 1, SG_VEL = 2, SG_OM = 3, SG_LV = 4, SG_LR = 5, SG_C0 = 6, SG_TG = 7, SG_C1 = 8 };
template <int MODE>
__host__ __device__ constexpr int ds_stage_bytes() {
  return MODE != 1 ? (8 * 16 * DS_TILE + 8 * DS_TILE + 16) : (5 * 16 * DS_TILE);
}

__device__ __forceinline__ uint32_t ds_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ds_mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ds_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ds_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ds_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ds_mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DS_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DS_DONE_%=;\n\t"
      "bra DS_WAIT_%=;\n\t"
      "DS_DONE_%=:\n\t}" ::"r"(ds_smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void ds_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ds_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ds_smem_u32(bar)) : "memory");
}
// bulk L2 prefetch (no shared-memory destination)
__device__ __forceinline__ void ds_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// the per-vehicle target array the mode stages (nullptr: none)
__device__ __forceinline__ const float4* ds_staged_target(const DsArgs& a) {
  return a.tmode == 0 ? a.t_pos : a.tmode == 2 ? a.t_vel : a.tmode == 3 ? a.rate_thrust : a.t_off;
}

// issued by ONE thread: all bulk copies of `tile` into stage memory `st`
template <bool NU6, int MODE>
__device__ __forceinline__ void ds_stage_issue(const DsArgs& a, int tile, unsigned char* st, unsigned long long* bar) {
  const long long v0 = (long long)tile * a.tile_v;
  const int cnt = (int)min((long long)a.tile_v, (long long)a.n - v0);  // vehicles of the tile that exist (> 0)
  const uint32_t b16 = (uint32_t)cnt * 16u;
  constexpr uint32_t ROW = 16u * DS_TILE;
  const float4* tg = (MODE != 1) ? ds_staged_target(a) : nullptr;
  // float2 array: keep the slab 16-byte aligned and inside the allocation (n_pad is even)
  const long long v0e = v0 & ~1LL;
  const uint32_t b8 = (MODE != 1 && NU6) ? (uint32_t)((((cnt + (int)(v0 - v0e)) * 8) + 15) & ~15) : 0u;
  const uint32_t total = (MODE != 1 ? 7u : 5u) * b16 + (tg ? b16 : 0u) + b8;
  ds_mbar_expect_tx(bar, total);
  ds_bulk_g2s(st + SG_POS * ROW, a.s_pos + v0, b16, bar);
  ds_bulk_g2s(st + SG_QUAT * ROW, a.s_quat + v0, b16, bar);
  ds_bulk_g2s(st + SG_VEL * ROW, a.s_vel + v0, b16, bar);
  ds_bulk_g2s(st + SG_OM * ROW, a.s_om + v0, b16, bar);
  ds_bulk_g2s(st + SG_LV * ROW, a.s_lv + v0, b16, bar);
  if (MODE != 1) {
    ds_bulk_g2s(st + SG_LR * ROW, a.s_lr + v0, b16, bar);
    ds_bulk_g2s(st + SG_C0 * ROW, a.s_c0 + v0, b16, bar);
    if (tg) ds_bulk_g2s(st + SG_TG * ROW, tg + v0, b16, bar);
    if (NU6) ds_bulk_g2s(st + SG_C1 * ROW, a.s_c1 + v0e, b8, bar);
    if (a.tmode == 0) {
      if (a.t_vel) ds_prefetch_l2(a.t_vel + v0, b16);
      if (a.t_acc) ds_prefetch_l2(a.t_acc + v0, b16);
    }
  }
}

// tg0: this vehicle's row of the mode's per-vehicle target array (staged copy in the step kernel, global otherwise)
__device__ __forceinline__ CtrlTarget ds_fetch_target(const DsArgs& a, const DsTypeDev& tp, const CtrlState& cs, int v,
                                                      int& wp, const float4* tg0) {
  CtrlTarget t;
  if (a.tmode == 2) {  // VelocityAviary._preprocessAction (VelocityAviary.py:236-257)
    const float4 q = *tg0;
    float n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    float k = (n2 != 0.f) ? tp.speed_limit * fabsf(q.w) / sqrtf(n2) : 0.f;
    float roll, pitch, yaw;
    ds_euler<true>(cs.qx, cs.qy, cs.qz, cs.qw, roll, pitch, yaw);  // the controller's (libm-free) variant
    t.x = cs.px; t.y = cs.py; t.z = cs.pz; t.yaw = yaw;  // hold position and yaw (state[0:3], state[9])
    t.vx = k * q.x; t.vy = k * q.y; t.vz = k * q.z;
    t.ax = t.ay = t.az = 0.f;
  } else if (a.tmode == 0) {
    const float4 p = *tg0;
    t.x = p.x; t.y = p.y; t.z = p.z; t.yaw = p.w;
    t.vx = t.vy = t.vz = t.ax = t.ay = t.az = 0.f;
    if (a.t_vel) { float4 q = __ldg(a.t_vel + v); t.vx = q.x; t.vy = q.y; t.vz = q.z; }
    if (a.t_acc) { float4 q = __ldg(a.t_acc + v); t.ax = q.x; t.ay = q.y; t.az = q.z; }
  } else {
    // the caller's index for this step (fly_INDI.py:230-245), else the resident counter; either is clamped: an index left
    // over from a longer table must not read past this one
    int w = a.t_wp ? __ldg(a.t_wp + v) : wp;
    w = min(max(w, 0), a.num_wp - 1);
    const float4* row = a.t_table + 3 * w;
    float4 p = __ldg(row), q = __ldg(row + 1), r = __ldg(row + 2);
    t.x = p.x; t.y = p.y; t.z = p.z; t.yaw = p.w;
    t.vx = q.x; t.vy = q.y; t.vz = q.z; t.ax = r.x; t.ay = r.y; t.az = r.z;
    if (a.t_off) { const float4 o = *tg0; t.x += o.x; t.y += o.y; t.z += o.z; }
    if (!a.t_wp) wp = a.advance_wp ? ((w < a.num_wp - 1) ? w + 1 : 0) : w;  // fly_INDI.py:242-245
  }
  return t;
}

// HOMO: every slot flies the same airframe type.  The type table then travels in the kernel parameters (constant bank:
// its entries are uniform operands of the arithmetic instructions, no LDS, no shared-memory copy) and the control law
// branch is warp-uniform.  Mixed swarms (HOMO = false) index the per-type tables in shared memory by the lane's type.
struct DsArgsH { DsArgs a; DsTypeDev tp; };
template <bool HOMO> struct DsKArgs { typedef DsArgs type; };
template <> struct DsKArgs<true> { typedef DsArgsH type; };
__device__ __forceinline__ const DsArgs& ds_args_of(const DsArgs& A) { return A; }
__device__ __forceinline__ const DsArgs& ds_args_of(const DsArgsH& A) { return A.a; }
__device__ __forceinline__ const DsTypeDev& ds_type_of(const DsArgs&, const DsTypeDev* sh, int id) { return sh[id]; }
__device__ __forceinline__ const DsTypeDev& ds_type_of(const DsArgsH& A, const DsTypeDev*, int) { return A.tp; }

template <int INTEG, int DW, bool NU6, bool WARPSYNC, int MODE, int FX, bool EXT, bool HOMO, bool RC>
__global__ void __launch_bounds__(DS_TILE, DS_MIN_CTAS) ds_step_kernel(const __grid_constant__ typename DsKArgs<HOMO>::type A) {
  const DsArgs& a = ds_args_of(A);
  extern __shared__ __align__(128) unsigned char ds_stage_mem[];  // 2 x ds_stage_bytes<MODE>()
  __shared__ __align__(8) unsigned long long sh_bar[2];  // per stage: the stage's bulk copies have landed
  __shared__ int sh_tile[2];                             // per stage: the tile staged there, -1 = none (the CTA is done)
  __shared__ __align__(16) DsTypeDev sh_types[HOMO ? 1 : DS_MAX_TYPES_DEV];
  __shared__ uint8_t sh_slot_type[32];
  __shared__ __align__(16) float4 sh_pos[DW ? 2 * DS_DW_BUF : 1];
  __shared__ float sh_stat[ST_COUNT * DS_TILE];
  constexpr int STAGE = ds_stage_bytes<MODE>();
  constexpr uint32_t ROW = 16u * DS_TILE;
  if (threadIdx.x == 0) {
    ds_mbar_init(&sh_bar[0], 1);
    ds_mbar_init(&sh_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the CTA's first two tiles (static) start streaming in while the type tables are being loaded; every further
    // tile comes from the atomic counter, so that no SM idles while another still has a whole tile queued: -4.3 % on
    // the K = 8 mixed swarm (more than the 1.2 % of static imbalance: SMs do not run at the same pace), +2 % on the
    // K = 2 workloads whose tiles last 3 us; choosing static / dynamic per launch was measured and lost on both.
    sh_tile[0] = blockIdx.x;
    ds_stage_issue<NU6, MODE>(a, blockIdx.x, ds_stage_mem, &sh_bar[0]);
    const int t1 = blockIdx.x + gridDim.x;
    sh_tile[1] = (t1 < a.n_tiles) ? t1 : -1;
    if (t1 < a.n_tiles) ds_stage_issue<NU6, MODE>(a, t1, ds_stage_mem + STAGE, &sh_bar[1]);
  }
  if (!HOMO) ds_load_types(a, sh_types);
  if (!HOMO && threadIdx.x < 32) sh_slot_type[threadIdx.x] = (threadIdx.x < a.D) ? a.slot_type[threadIdx.x] : 0;
  const bool stats_on = (a.flags & 8u) != 0;
  if (stats_on) {
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) sh_stat[i * DS_TILE + threadIdx.x] = (i == ST_MINZ) ? 3.0e38f : 0.f;
  }
  __syncthreads();

  constexpr int NU = NU6 ? 6 : 4;
  const int tid = threadIdx.x;
  const bool lane_ok = tid < a.tile_v;
  const int lv = lane_ok ? tid : 0;  // idle lanes shadow local vehicle 0 (no stores) so barriers stay uniform
  const int slot = lv % a.D;
  // row of the env's slot 0 in the downwash snapshot: D + 1 padded rows per env, or a 32-row block (symmetric variant)
  const int env_row0 = (DW == 2) ? (lv / 16) * DS_DW_SYM_ROWS : (lv / a.D) * (a.D + DS_DW_PAD);
  const int my_row = env_row0 + slot;
  const int type_id = HOMO ? a.homo_type : sh_slot_type[slot];
  const DsTypeDev& tp = ds_type_of(A, sh_types, type_id);

  for (int iter = 0;; ++iter) {
    const int tile = sh_tile[iter & 1];
    if (tile < 0) break;
    // thread 0 draws the tile this stage will hold two iterations from now; the atomic's latency hides behind the tile
    // (the raw ticket is not touched before the end of the tile: its first use is where the warp waits for the atomic)
    int ticket = 0;
    if (tid == 0) ticket = atomicAdd(a.tile_counter, 1);
    const int v = tile * a.tile_v + lv;
    const bool valid = lane_ok && v < a.n;
    const int vv = valid ? v : 0;
    const int ld = valid ? lv : 0;  // staged row this lane reads (vehicles past the end shadow the tile's first one)
    const unsigned char* st = ds_stage_mem + (iter & 1) * STAGE;
    ds_mbar_wait(&sh_bar[iter & 1], (uint32_t)(iter >> 1) & 1u);
    const float4* sg = reinterpret_cast<const float4*>(st);
    constexpr int T = DS_TILE;

    // ---- physics inputs.  The controller memory (last_vel, last_rates, cmd) is read from the stage only when the
    // control law runs, so that it does not occupy registers during the K substeps.
    const float4 P = sg[SG_POS * T + ld], Q = sg[SG_QUAT * T + ld], V = sg[SG_VEL * T + ld], W = sg[SG_OM * T + ld];
    PhysState s = {P.x, P.y, P.z, Q.x, Q.y, Q.z, Q.w, V.x, V.y, V.z, W.x, W.y, W.z};
    float prev_rpm_sum = V.w;
    float lthrust = P.w;
    int wp = __float_as_int(W.w);

    CtrlMem m;
    CtrlOut o = {0.f, 0.f, 0.f, 0.f, 0, 0};
    float perr = 0.f;
    uint32_t done_bits = 0;
    auto control = [&]() {  // INDIControl.computeControl on the resident state
      const float4 LV = sg[SG_LV * T + ld], LR = sg[SG_LR * T + ld], C0 = sg[SG_C0 * T + ld];
      m.lvx = LV.x; m.lvy = LV.y; m.lvz = LV.z; m.lrx = LR.x; m.lry = LR.y; m.lrz = LR.z; m.lthrust = lthrust;
      m.cmd[0] = C0.x; m.cmd[1] = C0.y; m.cmd[2] = C0.z; m.cmd[3] = C0.w;
      if (NU6) {
        const int lead = (tile * a.tile_v) & 1;  // the float2 slab starts at an even vehicle index
        const float2 C1 = reinterpret_cast<const float2*>(st + SG_C1 * ROW)[lead + ld];
        m.cmd[4] = C1.x; m.cmd[5] = C1.y;
      } else {
        m.cmd[4] = m.cmd[5] = 0.f;
      }
      done_bits = __float_as_uint(LV.w) & ~DS_PENDING_ACTION;
      if (EXT) { const float4 AF = a.s_af[vv]; m.afx = AF.x; m.afy = AF.y; m.afz = AF.z; }
      CtrlState cs = {s.px, s.py, s.pz, s.qx, s.qy, s.qz, s.qw, s.vx, s.vy, s.vz, s.wx, s.wy, s.wz};
      const float4* tg0 = sg + SG_TG * T + ld;
      if (a.tmode == 3) {  // RPYTAviary._preprocessAction -> INDIControl._INDIRateControl (RPYTAviary.py:180-193)
        const float4 rt = *tg0;
        float nu[4];
        ds_rate_loop<EXT>(tp, cs, a.inv_ctrl_dt, a.acc_b, rt.x, rt.y, rt.z, m, nu);
        nu[3] = rt.w - m.lthrust;  // INDIControl.py:454
        m.lthrust = rt.w;
        ds_allocate_quad<NU6>(tp, nu, m, o);
      } else {
        CtrlTarget t = ds_fetch_target(a, tp, cs, vv, wp, tg0);
        // the fused kernel never runs the FP64 active-set loop itself (order 1 with 6-DOF types is un-fused by the host)
        const WlsQueue wq = {a.wls_count, a.wls_index, a.wls_nu, valid ? v : -1};
        ds_indi_control<NU6, EXT, true>(tp, a.wls, type_id, cs, t, a.inv_ctrl_dt, a.acc_b, m, o, false, &wq);
      }
      perr = sqrtf(o.pex * o.pex + o.pey * o.pey + o.pez * o.pez);
      lthrust = m.lthrust;
    };

    // ---- the action the physics applies
    float act[6];
    if (MODE == 2) {  // VelocityAviary order: control, then physics with the new command
      control();
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = m.cmd[i];
    } else if (MODE == 1) {  // external action, clipped (CtrlAviary.py:258-263)
      const float* ea = a.ext_action + (size_t)vv * 6;
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = ds_clampf(ea[i], tp.rotor[i].pmin, tp.rotor[i].pmax);
    } else if (a.use_act || (__float_as_uint(sg[SG_LV * T + ld].w) & DS_PENDING_ACTION)) {
      // first step after a reset (of the whole batch, or of this vehicle's env): the caller's initial action (fly_INDI.py:214)
      const float4 A0 = a.s_a0[vv];
      act[0] = A0.x; act[1] = A0.y; act[2] = A0.z; act[3] = A0.w;
      if (NU6) { const float2 A1 = a.s_a1[vv]; act[4] = A1.x; act[5] = A1.y; }
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = ds_clampf(act[i], tp.rotor[i].pmin, tp.rotor[i].pmax);
    } else {  // the resident controller command, already clipped by the controller (INDIControl.py:487)
      const float4 C0 = sg[SG_C0 * T + ld];
      act[0] = C0.x; act[1] = C0.y; act[2] = C0.z; act[3] = C0.w;
      if (NU6) {
        const int lead = (tile * a.tile_v) & 1;
        const float2 C1 = reinterpret_cast<const float2*>(st + SG_C1 * ROW)[lead + ld];
        act[4] = C1.x; act[5] = C1.y;
      }
    }

    // table targets: the row this vehicle's control law will read after the K substeps is requested into L1 now (its
    // index sits in the staged tile), so that the gather's latency hides behind the physics - the K = 2 table workloads
    // spent 16 % of their stall samples waiting for it.  Single-vehicle envs only: in the downwash kernels the K substeps are
    // long enough to make the gather irrelevant, and the extra branch cost the 16-drone swarm 1 % (code generation).
    if (MODE == 0 && DW == 0 && a.tmode == 1 && !a.t_wp) {
      const float4* row = a.t_table + 3 * min(max(wp, 0), a.num_wp - 1);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 2));
    }

    float rpm[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // actual rotor speeds (motor model, EXT only)
    if (EXT) {
      const float4 R0 = a.s_r0[vv];
      rpm[0] = R0.x; rpm[1] = R0.y; rpm[2] = R0.z; rpm[3] = R0.w;
      if (NU6) { const float2 R1 = a.s_r1[vv]; rpm[4] = R1.x; rpm[5] = R1.y; }
    }
    ds_physics<INTEG, DW, NU6, WARPSYNC, FX, EXT, RC>(a, tp, env_row0, my_row, sh_pos, act, s, prev_rpm_sum, rpm, a.veh0 + (uint32_t)vv);
    if (MODE == 0) control();
    if (MODE == 1) done_bits = __float_as_uint(sg[SG_LV * T + ld].w) & ~DS_PENDING_ACTION;

    // ---- done predicate on the fresh state (fly_INDI_TrajectoryTrack.py:249-250)
    // |pos - goal|^2 < r^2 in a fixed order of correctly rounded operations (no contraction, no square root), so the bit
    // is a pure function of the FP32 position, reproduced exactly by the same float32 expression
    if (a.goal_en) {
      const float dx = s.px - a.goal_x, dy = s.py - a.goal_y, dz = s.pz - a.goal_z;
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 < a.goal_r2) done_bits |= 1u;
    }
    if (a.floor_en && s.pz < a.z_min) done_bits |= 2u;
    if (a.env_t0 ? (a.max_steps > 0 && a.step_end - __ldg(a.env_t0 + vv / a.D) >= a.max_steps) : (a.time_hit != 0)) done_bits |= 4u;

    // ---- per-env done / reward (CtrlAviary._computeDone / _computeReward, CtrlAviary.py:267-293, batched): a butterfly of
    // warp shuffles over the D lanes of the env, the env's slot 0 writes.  Same rule as ds_obs_kernel: the goal bit counts
    // for slot 0 only (the example tests drone "0"), floor / time for every slot.
    if (WARPSYNC && MODE != 1 && (a.env_done != nullptr || a.env_reward != nullptr)) {
      uint32_t db = valid ? ((slot == 0) ? done_bits : (done_bits & 6u)) : 0u;
      float pe = valid ? perr : 0.f;
      for (int o = a.D >> 1; o > 0; o >>= 1) {
        db |= __shfl_xor_sync(0xffffffffu, db, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
      }
      if (valid && slot == 0) {
        const int env = v / a.D;
        if (a.env_done) a.env_done[env] = db ? 1 : 0;
        if (a.env_reward) a.env_reward[env] = (a.reward_mode == 1) ? -pe / (float)a.D : -1.0f;
      }
    }

    if (valid) {
      a.s_pos[v] = make_float4(s.px, s.py, s.pz, lthrust);
      a.s_quat[v] = make_float4(s.qx, s.qy, s.qz, s.qw);
      a.s_vel[v] = make_float4(s.vx, s.vy, s.vz, prev_rpm_sum);
      a.s_om[v] = make_float4(s.wx, s.wy, s.wz, __int_as_float(wp));
      if (EXT) {
        a.s_r0[v] = make_float4(rpm[0], rpm[1], rpm[2], rpm[3]);
        if (NU6) a.s_r1[v] = make_float2(rpm[4], rpm[5]);
        if (MODE != 1) a.s_af[v] = make_float4(m.afx, m.afy, m.afz, 0.f);
      }
      if (MODE != 1) {
        a.s_lv[v] = make_float4(m.lvx, m.lvy, m.lvz, __uint_as_float(done_bits));
        a.s_lr[v] = make_float4(m.lrx, m.lry, m.lrz, perr);
        a.s_c0[v] = make_float4(m.cmd[0], m.cmd[1], m.cmd[2], m.cmd[3]);
        if (NU6) a.s_c1[v] = make_float2(m.cmd[4], m.cmd[5]);
      } else {
        reinterpret_cast<float*>(a.s_lv + v)[3] = __uint_as_float(done_bits);
      }
      if (a.store_act) {
        a.s_a0[v] = make_float4(act[0], act[1], act[2], act[3]);
        if (NU6) a.s_a1[v] = make_float2(act[4], act[5]);
      }
      if (stats_on) {
        float* sc = sh_stat + tid;
        if (MODE != 1) {
          sc[ST_NCTRL * DS_TILE] += 1.f;
          sc[ST_ERR2 * DS_TILE] += perr * perr;
          if (o.sat) sc[ST_SAT * DS_TILE] += (float)o.sat;
          if (o.wls_iter != 1 && o.wls_iter != 0) sc[ST_WLS_SLOW * DS_TILE] += 1.f;
          if (o.wls_iter < 0) sc[ST_WLS_FAIL * DS_TILE] += 1.f;
        }
        const bool fin = isfinite(s.px) && isfinite(s.py) && isfinite(s.pz) && isfinite(s.qw) && isfinite(s.vx) && isfinite(s.wx);
        if (!fin) sc[ST_NONFINITE * DS_TILE] += 1.f;
        sc[ST_MINZ * DS_TILE] = fminf(sc[ST_MINZ * DS_TILE], s.pz);
        if (done_bits) sc[ST_DONE * DS_TILE] += 1.f;
      }
    }
    // ---- every thread has read its rows of this stage: refill it with the tile two iterations ahead.  A CTA-wide
    // barrier per tile, on purpose: releasing the stage per warp (an "empty" mbarrier that only thread 0 waits on) was
    // measured 4 % slower - the barrier keeps the CTA's warps in the same region of the 9000-instruction kernel
    // (instruction-cache locality), and the waiting thread's try_wait loop competes for issue slots.
    __syncthreads();
    if (tid == 0) {
      const int next_tile = 2 * (int)gridDim.x + ticket;
      const bool more = next_tile < a.n_tiles;
      // every processed tile draws exactly one ticket and grid <= n_tiles, so a launch draws n_tiles tickets: whoever holds
      // the last one re-arms the counter for the next launch (nobody draws after it), with no end-of-kernel handshake
      if (ticket == a.n_tiles - 1) *a.tile_counter = 0;
      sh_tile[iter & 1] = more ? next_tile : -1;  // read two iterations (two barriers) from now
      if (more) ds_stage_issue<NU6, MODE>(a, next_tile, ds_stage_mem + (iter & 1) * STAGE, &sh_bar[iter & 1]);
    }
  }
  if (stats_on) {
    if (DW == 0) ds_flush_stats(sh_stat, a.stats);
    else ds_flush_stats_call(sh_stat, a.stats);
  }
}
