// One-warp-per-environment fused step kernel for N <= 64 pedestrians, float32 (the BASELINE headline
// shapes: 60 pedestrians).  Same semantics as evac_step_kernel (evac_kernels.cuh; reference file:line
// citations there and inline below) but written as straight-line warp code:
//   * lane L owns pedestrians L and L + 32; every per-pedestrian quantity is kept as an (x, y) float2 so the
//     epilogue arithmetic runs on Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2) exactly like the
//     pairwise pass -- one instruction per pedestrian instead of one per component;
//   * the observation encoding is a template parameter (no per-step mode branches), all loads of a step are
//     issued up front (incl. the read-modify-write words of lane 0), reductions are one REDUX + one shuffle tree;
//   * the angular noise is one Philox2x32-10 block per lane (two words = the lane's two pedestrians), no
//     shared-memory staging; no block barrier anywhere (one warp: __syncwarp only).
#pragma once
#include "evac_kernels.cuh"

namespace evac {

enum { WMODE_GENERIC = 0, WMODE_REL_OHE_BOX = 1, WMODE_GRAV = 2 };

__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// Neighbour sums of ONE pedestrian over the slot pairs [j0, j1) of the strip-sorted tile (same packed evaluation as
// pairwise_pass; trip counts differ per lane, finished lanes idle).
__device__ __forceinline__ void windowed_pass(const Tile<float>& t, int j0, int j1, float x, float y, float thr2, float& sx, float& sy) {
  const float2 nx = splat(-x), ny = splat(-y);
  float2 ax = make_float2(0.f, 0.f), ay = make_float2(0.f, 0.f);
  if (j0 < j1) {
    float4 p = t.P2[j0], u = t.U2[j0];
#pragma unroll 2
    for (int j = j0; j < j1; ++j) {
      const float4 pn = t.P2[j + 1], un = t.U2[j + 1];  // look-ahead (at most the tile's spare entry)
      const float2 dx = __fadd2_rn(make_float2(p.x, p.y), nx);
      const float2 dy = __fadd2_rn(make_float2(p.z, p.w), ny);
      float2 d2 = __fmul2_rn(dx, dx);
      d2 = __ffma2_rn(dy, dy, d2);
      const float2 w = make_float2(d2.x < thr2 ? 1.f : 0.f, d2.y < thr2 ? 1.f : 0.f);
      ax = __ffma2_rn(w, make_float2(u.x, u.y), ax);
      ay = __ffma2_rn(w, make_float2(u.z, u.w), ay);
      p = pn;
      u = un;
    }
  }
  sx = ax.x + ax.y;
  sy = ay.x + ay.y;
}

// per-pedestrian working set of the warp kernel
struct WPed {
  float2 p, d;  // position, direction
  int st;       // status (ST_NONE for padding lanes)
};

#ifndef EVAC_WARP_MINB
// resident warps per SM the register allocation is sized for.  28 -> 70 registers per thread, no spills in the step loop, and
// still 29 warps per SM: the headline batch (4096 envs = 27.7 warps per SM) stays one wave.  Measured on B200 against 32
// (64 registers, 8 bytes spilled per step): 9.23 -> 8.90 us per 4096-env step, 5.11 -> 4.92 us in-rollout; 24 (76 registers,
// 26 warps per SM) pushes 4096 envs into a second wave: 9.81 us.
#define EVAC_WARP_MINB 28
#endif
#ifndef EVAC_PAIR_UNROLL
#define EVAC_PAIR_UNROLL 4  // slot pairs per unrolled iteration of the pairwise loop
#endif
// WPC = environments (independent warps) per CTA: fewer, fatter CTAs for the block scheduler; no block-level barrier anywhere.
template <int MODE, int WPC>
__global__ void __launch_bounds__(32 * WPC, EVAC_WARP_MINB / WPC) evac_warp_kernel(const __grid_constant__ KArgs<float> a) {  // @region wload
  // Tile<float> of 64 slots (+ the look-ahead entries) per warp = 66 float4.  EVAC_OBS_STAGED (A/B variant): after the pairwise
  // pass the same memory stages the observation row ((64 + 2) x 6 floats = 99 float4), see the observation section
#if defined(EVAC_OBS_STAGED) || defined(EVAC_OBS_BULK)
  __shared__ __align__(128) float4 tile_all[WPC][99];
#else
  __shared__ __align__(16) float4 tile_all[WPC][66];
#endif
  // strip culling (a.cells_x > 0): the sources are sorted by vertical strip (edge >= vision radius) so that a
  // pedestrian only visits the slots of its own and the two adjacent strips
  __shared__ uint2 strip_mask_all[WPC][32];   // per strip: which lanes' pedestrian 0 (.x) / pedestrian 1 (.y) sit in it
  __shared__ int strip_start_all[WPC][33];    // first slot of every strip (+ total)
  const int wic = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31, e = blockIdx.x * WPC + wic, N = a.N;
  if (e >= a.E) return;
#ifdef EVAC_PROBE_EMPTY  // measurement variant (never shipped): launch + CTA scheduling cost only
  if (a.N >= 0) return;
#endif
  const Tile<float> tile(reinterpret_cast<unsigned char*>(tile_all[wic]), 64);
  uint2* strip_mask = strip_mask_all[wic];
  int* strip_start = strip_start_all[wic];
  const int S = a.cells_x;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const bool valid[2] = {lane < N, lane + 32 < N};
  const uint32_t env_g = (uint32_t)(a.env_offset + e);
  const float2 exit_p = make_float2(0.f, -1.f);  // area.py:36-39

  // ---------------- load state: everything this launch reads is requested before anything is used.  One address per
  // environment (EnvBlock, evac_kernels.cuh): the record as two broadcast LDG.128, one LDG.128 per pedestrian
  // ({px, py, dx, dy}), one LDG.U16 for the lane's two statuses; slots >= N hold zeros / status 0 and are never valid.
  unsigned char* const blk = a.blocks + (size_t)e * BLK_BYTES;
  float4* const ped_l = reinterpret_cast<float4*>(blk + BLK_PED) + lane;
  unsigned short* const st_l = reinterpret_cast<unsigned short*>(blk + BLK_STATUS) + lane;
  const float4 pd0 = ped_l[0], pd1 = ped_l[32];
  const unsigned st2 = *st_l;
  const float4 rec0 = *reinterpret_cast<const float4*>(blk);      // agent_pos, agent_dir
  const int4 rec1 = *reinterpret_cast<const int4*>(blk + 16);     // now, episode, agent_state
  // overall_timesteps (all lanes: the RotatingAgent's call counter) and, lane 0, the episode accumulators of env.py:168-170
  const longlong2 rec2 = *reinterpret_cast<const longlong2*>(blk + 32);
  long long overall = rec2.x;
  double acc_r = __longlong_as_double(rec2.y), acc_i = 0, acc_s = 0;
  if (lane == 0) {
    const double2 r3 = *reinterpret_cast<const double2*>(blk + 48);
    acc_i = r3.x; acc_s = r3.y;
  }
  WPed q[2];
  q[0].p = make_float2(pd0.x, pd0.y); q[0].d = make_float2(pd0.z, pd0.w); q[0].st = (int)(st2 & 0xffu);
  q[1].p = make_float2(pd1.x, pd1.y); q[1].d = make_float2(pd1.z, pd1.w); q[1].st = (int)(st2 >> 8);
  float2 ap = make_float2(rec0.x, rec0.y), ad = make_float2(rec0.z, rec0.w);
  int now = rec1.x, episode = rec1.y, wac_state = rec1.z;
  float reward_sum = 0.f;
  int any_term = 0, any_trunc = 0;
  const float noise_c = a.noise_coef;
  const float* noise_e = a.noise ? a.noise + (size_t)e * N : nullptr;
  float* obs_e = a.obs ? a.obs + (size_t)e * a.obs_dim : nullptr;
  const float2 wall = make_float2(a.width, a.height);

  for (int s = 0; s < a.num_steps; ++s) {  // @region wrng
#ifdef EVAC_PROBE_FLOOR  // measurement variant (never shipped): load -> observation -> write back, no dynamics
    float2 act_tbl = make_float2(0.f, 0.f);
    if (a.agent_kind == AGENT_TABLE) act_tbl = a.actions[(size_t)s * a.E + e];
    ap.x += 1e-9f * act_tbl.x;
    reward_sum += act_tbl.y;
#else
    // ---------------- Time.step [area.py:53-59]
    const int now_prev = now;
    now += 1;
    const bool truncated = now >= a.max_timesteps;
    // the action is requested first so that its (cold) latency overlaps the preparation below
    float2 act_tbl = make_float2(0.f, 0.f);
    if (a.agent_kind == AGENT_TABLE) act_tbl = a.actions[(size_t)s * a.E + e];
    // ---------------- angular noise of this lane's two pedestrians [area.py:124]
    float nz[2];
    if (noise_e != nullptr) {
      const float* np_ = noise_e + (size_t)s * a.E * N;
      nz[0] = valid[0] ? np_[lane] : 0.f;
      nz[1] = valid[1] ? np_[lane + 32] : 0.f;
    } else {
      const uint2 r = evac_noise_block(a.seed, env_g, (uint32_t)episode, (uint32_t)now_prev, (uint32_t)lane);
      nz[0] = (u01(r.x) - 0.5f) * noise_c;
      nz[1] = (u01(r.y) - 0.5f) * noise_c;
    }
    // ---------------- escaped / exiting preparation + unit directions [area.py:79-101]  // @region wprep
    float2 u[2];
    bool efv[2], fv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int so = q[k].st;
      if (so == ST_ESCAPED) { q[k].d = make_float2(0.f, 0.f); q[k].p = exit_p; }
      float2 v = q[k].d;
      if (so == ST_EXITING) v = __fadd2_rn(exit_p, make_float2(-q[k].p.x, -q[k].p.y));
      efv[k] = (unsigned)(so - ST_VISCEK) <= (unsigned)(ST_EXITING - ST_VISCEK);
      fv[k] = (unsigned)(so - ST_VISCEK) <= (unsigned)(ST_FOLLOWER - ST_VISCEK);
      // u = v / |v|; a zero direction gives NaN exactly like area.py:101
      const float2 sq = __fmul2_rn(v, v);
      const float n2 = sq.x + sq.y;
      const float inv = inv_norm(n2);
      u[k] = __fmul2_rn(v, splat(inv));
      if (so == ST_EXITING) q[k].d = __fmul2_rn(v, splat(fminf(n2 * inv, a.step_size) * inv));  // dir = u * min(|v|, step)
    }
    // ---------------- compact the moving pedestrians into the shared tile  // @region wcompact
    int n_src;
    int win_lo[2] = {0, 0}, win_hi[2] = {0, 0};  // slot window of this lane's pedestrians (strip culling)
    bool culled = false;
    if (S > 0) {
      // deterministic counting sort by strip: atomicOr of lane bits (commutative, so the result does not depend on
      // the order the atomics retire), rank = number of lower (pedestrian, lane) pairs of the same strip
      int c[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) c[k] = min(max((int)((q[k].p.x + a.width_f) * a.cell_inv_x), 0), S - 1);  // (int)NaN == 0
      strip_mask[lane] = make_uint2(0u, 0u);
      __syncwarp();
      if (efv[0]) atomicOr(&strip_mask[c[0]].x, 1u << lane);
      if (efv[1]) atomicOr(&strip_mask[c[1]].y, 1u << lane);
      __syncwarp();
      {
        const uint2 m = strip_mask[lane];
        const int cnt = (lane < S) ? __popc(m.x) + __popc(m.y) : 0;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        strip_start[lane] = inc - cnt;
        if (lane == 31) strip_start[32] = inc;
        n_src = __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (efv[k]) {
          const uint2 m = strip_mask[c[k]];
          const int slot = strip_start[c[k]] + (k == 0 ? __popc(m.x & lt_mask) : __popc(m.x) + __popc(m.y & lt_mask));
          tile.put(slot, q[k].p.x, q[k].p.y, u[k].x, u[k].y);
        }
        if (fv[k]) {
          win_lo[k] = strip_start[max(c[k] - 1, 0)] >> 1;                      // in slot PAIRS: a neighbouring slot that
          win_hi[k] = (strip_start[min(c[k] + 1, S - 1) + 1] + 1) >> 1;        // rides along fails the distance test
        }
      }
      // cost model: two windowed loops (one pedestrian each) against one all-pairs loop shared by both pedestrians;
      // a source without a direction (NaN) poisons every sum in the reference -> all-pairs path reproduces that
      const int wmax0 = __reduce_max_sync(0xffffffffu, win_hi[0] - win_lo[0]), wmax1 = __reduce_max_sync(0xffffffffu, win_hi[1] - win_lo[1]);
      const bool nan_src = (efv[0] && (u[0].x != u[0].x || u[0].y != u[0].y)) || (efv[1] && (u[1].x != u[1].x || u[1].y != u[1].y));
      culled = (wmax0 + wmax1) * 5 < ((n_src + 1) >> 1) * 8 && !__any_sync(0xffffffffu, nan_src);
    } else {
      const uint32_t m0 = __ballot_sync(0xffffffffu, efv[0]), m1 = __ballot_sync(0xffffffffu, efv[1]);
      const int c0 = __popc(m0);
      if (efv[0]) tile.put(__popc(m0 & lt_mask), q[0].p.x, q[0].p.y, u[0].x, u[0].y);
      if (efv[1]) tile.put(c0 + __popc(m1 & lt_mask), q[1].p.x, q[1].p.y, u[1].x, u[1].y);
      n_src = c0 + __popc(m1);
    }
    if (lane < 2) tile.put(n_src + lane, PARK, PARK, 0.f, 0.f);  // pad to an even count
    __syncwarp();
    // ---------------- action source + Area.agent_step [area.py:182-210], IEEE float32 like the reference  // @region wagent
    float r_agent = 0.f;
    bool term_agent = false;
    {
      float ax, ay;
      if (a.agent_kind == AGENT_TABLE) {
        ax = act_tbl.x; ay = act_tbl.y;
      } else if (a.agent_kind == AGENT_RANDOM) {  // RandomAgent: action_space.sample() ~ U[-1,1)^2 [random_agent.py:8-9]
        const uint2 r = evac_agent_block(a.seed, env_g, (uint32_t)episode, (uint32_t)now_prev);
        ax = 2.f * u01(r.x) - 1.f; ay = 2.f * u01(r.y) - 1.f;
      } else if (a.agent_kind == AGENT_WACUUM) {
        wacuum_act(ap.x, ap.y, wac_state, a, ax, ay);
      } else {  // RotatingAgent [rotating_agent.py:8-16]: i counts the agent's act() calls and never restarts with an episode
        const double ph = 0.05 * (double)(overall + 1);  // float64 like the reference
        ax = (float)sin(ph); ay = (float)cos(ph);
      }
      const float nrm = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay))), a.eps_f);
      ax = __fdiv_rn(ax, nrm); ay = __fdiv_rn(ay, nrm);
      ad.x = __fmul_rn(a.step_size_f, ax); ad.y = __fmul_rn(a.step_size_f, ay);
      const float ptx = __fadd_rn(ap.x, ad.x), pty = __fadd_rn(ap.y, ad.y);
      const bool collide = (ptx < -a.width_f) | (ptx > a.width_f) | (pty < -a.height_f) | (pty > a.height_f);
      if (!collide) { ap.x = ptx; ap.y = pty; }
      else { r_agent = -5.f; term_agent = a.term_wall != 0; }
    }
    // ---------------- pairwise alignment [area.py:105-119]  // @region wpairwise
    float sx[2], sy[2], cnt[2];
    if (culled) {
#pragma unroll
      for (int k = 0; k < 2; ++k) windowed_pass(tile, win_lo[k], win_hi[k], q[k].p.x, q[k].p.y, a.thr2_ped, sx[k], sy[k]);
    } else {
      const float xi[2] = {q[0].p.x, q[1].p.x}, yi[2] = {q[0].p.y, q[1].p.y};
#ifdef EVAC_PROBE_NOPAIR  // measurement variant (never shipped): the step without its pairwise pass
      sx[0] = sx[1] = q[0].d.x; sy[0] = sy[1] = q[1].d.y;
#else
      if (__any_sync(0xffffffffu, fv[0] | fv[1])) pairwise_pass<2, false, EVAC_PAIR_UNROLL>(tile, n_src, xi, yi, a.thr2_ped, sx, sy, cnt);
      else sx[0] = sx[1] = sy[0] = sy[1] = 0.f;
#endif
    }
    // ---------------- new headings, enslaving, integration, reflection, statuses  // @region wupdate
    const float2 e_ad = make_float2(__fmul_rn(a.enslaving_f, ad.x), __fmul_rn(a.enslaving_f, ad.y));  // float32 like area.py:140
    int counts = 0;  // k_exit | k_fol << 8 | n_esc << 16 | n_exi << 24   (each <= 64)
    float sum_dexit = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int so = q[k].st;
      if (fv[k]) {
        // cos/sin(atan2(my, mx) + nz) == unit(m) rotated by nz; atan2(0, 0) = 0 -> unit = (1, 0)  [area.py:120-136]
        const float m2 = fmaf(sx[k], sx[k], sy[k] * sy[k]);
        float2 c = make_float2(1.f, 0.f);
        if (m2 != 0.f) c = __fmul2_rn(make_float2(sx[k], sy[k]), splat(inv_norm(m2)));
        float sn, cn;
        sincos_noise(nz[k], sn, cn);
        float2 nd = __fmul2_rn(splat(c.x), make_float2(cn, sn));
        nd = __ffma2_rn(make_float2(-c.y, c.y), make_float2(sn, cn), nd);
        nd = __fmul2_rn(nd, splat(a.step_size));
        if (so == ST_FOLLOWER) nd = __ffma2_rn(splat(a.one_minus_enslaving), nd, e_ad);  // area.py:138-142
        q[k].d = nd;
      }
      if (efv[k]) q[k].p = __fadd2_rn(q[k].p, q[k].d);
      {  // wall reflection for ALL pedestrians [area.py:147-152]
        const float2 cl = make_float2(fminf(fmaxf(q[k].p.x, -wall.x), wall.x), fminf(fmaxf(q[k].p.y, -wall.y), wall.y));
        const float2 miss = __fadd2_rn(q[k].p, make_float2(-cl.x, -cl.y));
        q[k].p = __ffma2_rn(splat(-2.f), miss, q[k].p);
        if (miss.x != 0.f) q[k].d.x = -q[k].d.x;
        if (miss.y != 0.f) q[k].d.y = -q[k].d.y;
      }
      // statuses: a pure function of (position, agent position) [statuses.py:29-48]
      const float2 da = __fadd2_rn(make_float2(ap.x, ap.y), make_float2(-q[k].p.x, -q[k].p.y));
      const float2 de = __fadd2_rn(exit_p, make_float2(-q[k].p.x, -q[k].p.y));
      const float da2 = fmaf(da.x, da.x, da.y * da.y), de2 = fmaf(de.x, de.x, de.y * de.y);
      int sn_ = ST_VISCEK;
      if (da2 < a.thr2_leader) sn_ = ST_FOLLOWER;
      if (de2 < a.thr2_exit) sn_ = ST_EXITING;
      if (de2 < a.thr2_escape) sn_ = ST_ESCAPED;
      if (valid[k]) {
        sum_dexit += sqrt_fast(de2);
        counts += (int)(fv[k] && sn_ == ST_EXITING) + ((int)(so == ST_VISCEK && sn_ == ST_FOLLOWER) << 8) +
                  ((int)(sn_ == ST_ESCAPED) << 16) + ((int)(sn_ == ST_EXITING) << 24);
        q[k].st = sn_;
      }
    }
    // ---------------- warp reduction, rewards, termination [reward.py:19-46, area.py:174-180, env.py:158-171]  // @region wreward
    counts = __reduce_add_sync(0xffffffffu, counts);
    const float sd = warp_sum(sum_dexit);
    const int K_exit = counts & 0xff, K_fol = (counts >> 8) & 0xff, N_esc = (counts >> 16) & 0xff, N_exi = (counts >> 24) & 0xff;
    if (a.status_counts != nullptr) {  // pedestrians.status_stats after this step (before a same-step reset)
      const int n_fol = __popc(__ballot_sync(0xffffffffu, q[0].st == ST_FOLLOWER)) + __popc(__ballot_sync(0xffffffffu, q[1].st == ST_FOLLOWER));
      if (lane == 0)
        reinterpret_cast<ushort4*>(a.status_counts)[(size_t)s * a.E + e] =
            make_ushort4((unsigned short)N_esc, (unsigned short)N_exi, (unsigned short)n_fol, (unsigned short)(N - N_esc - N_exi - n_fol));
    }
    const float tf = 1.f - (float)now * a.inv_200n, intrinsic = 0.f - sd * a.inv_n;
    float r_ped = a.init_reward;
    if (a.exit_reward) r_ped += (15.f + 10.f * tf) * (float)K_exit;
    if (a.follow_reward) r_ped += (10.f + 5.f * tf) * (float)K_fol;
    const float r_status = r_agent + r_ped;
    const float reward = r_status + a.intrinsic_coef * intrinsic;
    const bool terminated = term_agent || (N_esc == N);
    reward_sum += reward;
    any_term |= terminated; any_trunc |= truncated;
    acc_r += (double)reward; acc_i += (double)intrinsic; acc_s += (double)r_status;
    overall += 1;
    // ---------------- same-step auto-reset (gymnasium vector-env semantics)  // @region wreset
    if (a.auto_reset && (terminated || truncated)) {
      const int N_fol = __popc(__ballot_sync(0xffffffffu, q[0].st == ST_FOLLOWER)) + __popc(__ballot_sync(0xffffffffu, q[1].st == ST_FOLLOWER));
      if (lane == 0) {  // the logging dict of env.py:115-125
        float* es = a.ep_stats + (size_t)e * NUM_EPISODE_STATS;
        const float v[NUM_EPISODE_STATS] = {(float)acc_i, (float)acc_s, (float)acc_r, (float)now, (float)N_esc, (float)N_exi,
                                            (float)N_fol, (float)(N - N_esc - N_exi - N_fol), (float)overall};
#pragma unroll
        for (int t = 0; t < NUM_EPISODE_STATS; ++t) { es[t] = v[t]; atomicAdd(a.totals + 1 + t, (double)v[t]); }
        atomicAdd(a.totals, 1.0);
        a.ep_finished[e] = 1;
      }
      acc_r = acc_i = acc_s = 0;
      now = 0; episode += 1; wac_state = 0;
      ap = make_float2(0.f, 0.f); ad = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (valid[k]) {
          float dummy;
          random_layout<float>(a.seed, env_g, (uint32_t)episode, (uint32_t)(lane + 32 * k), q[k].p.x, q[k].p.y, q[k].d.x, q[k].d.y);
          q[k].st = status_of<float>(q[k].p.x, q[k].p.y, 0.f, 0.f, a, dummy);
        }
      }
    }
#endif  // EVAC_PROBE_FLOOR
    // ---------------- observation [wrappers.py:8-96, gravity_encoding.py:8-81]  // @region wobs
#if defined(EVAC_PROBE_NOOBS)
    if (false) {
#else
    if (obs_e != nullptr && (a.obs_every_step || s == a.num_steps - 1)) {
#endif
      float* row = obs_e + (a.obs_every_step ? (size_t)s * a.E * a.obs_dim : (size_t)0);
      if constexpr (MODE == WMODE_REL_OHE_BOX) {
        // rows = [agent; exit; pedestrians], cols = [x, y, ohe(4)]; relative positions / sqrt(2) (float32 hypotenuse).
        // Default: three float2 stores per pedestrian straight from registers (24-byte row stride: a warp's STG.64 touches 24
        // sectors for 256 useful bytes, the sectors are completed in L2 by the neighbouring lanes' stores).
        // EVAC_OBS_STAGED: the row is assembled in shared memory and copied out with one 16-byte (N even) or 8-byte store per
        // lane and round -- 512 contiguous bytes per warp instruction.  Measured on B200, same box, 4096 x 60 per-step regime:
        // the load -> observe -> write-back skeleton gets faster (5.23 -> 4.82 us) but the FULL step slower (8.42 -> 8.66 us:
        // STS -> LDS -> STG adds a dependent shared-memory round trip to every warp's tail), so the direct stores stay.
        // EVAC_OBS_BULK: the staged row leaves as ONE cp.async.bulk (shared -> global, the TMA engine's 1-D path; UBLKCP in the
        // SASS) issued by lane 0, its wait deferred behind the state write-back: 8.76 vs 8.52 us, same box -- also slower.
        const float inv_hyp = (float)(1.0 / 1.41421353816986083984375);
        const float2 nap = make_float2(-ap.x, -ap.y);
#if !defined(EVAC_OBS_STAGED) && !defined(EVAC_OBS_BULK)
        if (lane == 0) {
          const float2 ex = __fmul2_rn(__fadd2_rn(exit_p, nap), splat(inv_hyp));
          float2* r = reinterpret_cast<float2*>(row);
          r[0] = ap; r[1] = make_float2(0.f, 0.f); r[2] = make_float2(0.f, 0.f);
          r[3] = ex; r[4] = make_float2(1.f, 0.f); r[5] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (valid[k]) {
            float2* r = reinterpret_cast<float2*>(row + (lane + 32 * k + 2) * 6);
            const int st = q[k].st;
            r[0] = __fmul2_rn(__fadd2_rn(q[k].p, nap), splat(inv_hyp));
            r[1] = make_float2(st == ST_ESCAPED ? 1.f : 0.f, st == ST_EXITING ? 1.f : 0.f);
            r[2] = make_float2(st == ST_FOLLOWER ? 1.f : 0.f, st == ST_VISCEK ? 1.f : 0.f);
          }
        }
#else
        float2* stage = reinterpret_cast<float2*>(tile_all[wic]);
        if (lane == 0) {
          const float2 ex = __fmul2_rn(__fadd2_rn(exit_p, nap), splat(inv_hyp));
          stage[0] = ap; stage[1] = make_float2(0.f, 0.f); stage[2] = make_float2(0.f, 0.f);  // agent row, status [0,0,0,0]
          stage[3] = ex; stage[4] = make_float2(1.f, 0.f); stage[5] = make_float2(0.f, 0.f);  // exit row,  status [1,0,0,0]
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (valid[k]) {
            float2* r = stage + (lane + 32 * k + 2) * 3;
            const int st = q[k].st;  // column 4 - status: ESCAPED -> 0, EXITING -> 1, FOLLOWER -> 2, VISCEK -> 3
            r[0] = __fmul2_rn(__fadd2_rn(q[k].p, nap), splat(inv_hyp));
            r[1] = make_float2(st == ST_ESCAPED ? 1.f : 0.f, st == ST_EXITING ? 1.f : 0.f);
            r[2] = make_float2(st == ST_FOLLOWER ? 1.f : 0.f, st == ST_VISCEK ? 1.f : 0.f);
          }
        }
#ifdef EVAC_OBS_BULK
        // A/B variant: the staged row leaves the SM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA
        // engine's 1-D path) issued by lane 0, N even (24 (N + 2) bytes is then a multiple of 16)
        if ((N & 1) == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
          __syncwarp();
          if (lane == 0) {
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile_all[wic]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row), "r"(src), "r"((N + 2) * 24) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the tile may only be rewritten (next step) once the copy has read it; after the last step the wait moves behind
            // the state write-back at the end of the kernel
            if (s != a.num_steps - 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
        } else {
          __syncwarp();
          const int n8 = (N + 2) * 3;
          float2* dst = reinterpret_cast<float2*>(row);
#pragma unroll
          for (int c = 0; c < 7; ++c) { const int j = lane + 32 * c; if (j < n8) dst[j] = stage[j]; }
        }
#else
        __syncwarp();
        if ((N & 1) == 0) {
          const int n16 = (N + 2) * 3 / 2;  // float4 chunks of the row
          const float4* src = tile_all[wic];
          float4* dst = reinterpret_cast<float4*>(row);
#pragma unroll
          for (int c = 0; c < 4; ++c) { const int j = lane + 32 * c; if (j < n16) dst[j] = src[j]; }
        } else {
          const int n8 = (N + 2) * 3;
          float2* dst = reinterpret_cast<float2*>(row);
#pragma unroll
          for (int c = 0; c < 7; ++c) { const int j = lane + 32 * c; if (j < n8) dst[j] = stage[j]; }
        }
#endif
#endif
      } else if (MODE == WMODE_GRAV || a.positions == POS_GRAV) {
        // gravity encoding [gravity_encoding.py:8-38]: per VISCEK pedestrian -alpha / (|R| + eps)^(alpha + 2) * R, R = agent - p.
        // Straight-line and unconditional for both pedestrians of the lane (select, not branch); MUFU square root and
        // reciprocal (~2 ulp each against the parity bar of 2e-5 of the largest term); the two sums over the warp in double.
        float gx = 0.f, gy = 0.f;
        int nf = 0;
        const float alpha = (float)a.alpha, eps = (float)a.eps;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float2 R = __fadd2_rn(ap, make_float2(-q[k].p.x, -q[k].p.y));
          const float norm = sqrt_fast(fmaf(R.x, R.x, R.y * R.y)) + eps;
          const float n2 = norm * norm;
          float pw;
          switch (a.alpha_plus2_int) {  // alpha = 2 .. 5 (BASELINE config 3) without a loop
            case 4: pw = n2 * n2; break;
            case 5: pw = n2 * n2 * norm; break;
            case 6: pw = n2 * n2 * n2; break;
            case 7: pw = n2 * n2 * n2 * norm; break;
            default: pw = a.alpha_plus2_int ? ipow(norm, a.alpha_plus2_int) : powf(norm, alpha + 2.f);
          }
          const float c = __fdividef(-alpha, pw);
          const bool vk = q[k].st == ST_VISCEK;
          gx += vk ? c * R.x : 0.f;
          gy += vk ? c * R.y : 0.f;
          nf += (q[k].st == ST_FOLLOWER);
        }
        nf = __reduce_add_sync(0xffffffffu, nf);
        const double wx = warp_sum((double)gx), wy = warp_sum((double)gy);
        if (lane == 0) store_grav_obs<float>(row, ap.x, ap.y, (float)wx, (float)wy, nf, a);
      } else {
        if (lane == 0) store_head_obs<float>(row, ap.x, ap.y, a);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (valid[k]) store_ped_obs<float>(row, lane + 32 * k, N, q[k].p.x, q[k].p.y, q[k].st, ap.x, ap.y, a);
      }
    }
    __syncwarp();  // the tile is rewritten by the next step
  }  // steps

  // ---------------- write back  // @region wwriteback
#if defined(EVAC_PROBE_NOSTATE)
  if (q[0].p.x == 123.456f)
#endif
  {
    ped_l[0] = make_float4(q[0].p.x, q[0].p.y, q[0].d.x, q[0].d.y);
    ped_l[32] = make_float4(q[1].p.x, q[1].p.y, q[1].d.x, q[1].d.y);
    *st_l = (unsigned short)(q[0].st | (q[1].st << 8));
  }
  if (a.status_out != nullptr) {  // host face: dense [E,N] statuses travel in the result block
    uint8_t* so = a.status_out + (size_t)e * N + lane;
    if (valid[0]) so[0] = (uint8_t)q[0].st;
    if (valid[1]) so[32] = (uint8_t)q[1].st;
  }
  if (lane == 0) {
    *reinterpret_cast<float4*>(blk) = make_float4(ap.x, ap.y, ad.x, ad.y);
    *reinterpret_cast<int4*>(blk + 16) = make_int4(now, episode, wac_state, 0);
    *reinterpret_cast<longlong2*>(blk + 32) = make_longlong2(overall, __double_as_longlong(acc_r));
    *reinterpret_cast<double2*>(blk + 48) = make_double2(acc_i, acc_s);
    if (a.reward) a.reward[e] = reward_sum;
    if (a.terminated) a.terminated[e] = (uint8_t)any_term;
    if (a.truncated) a.truncated[e] = (uint8_t)any_trunc;
#ifdef EVAC_OBS_BULK
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
  }
}

}  // namespace evac
