// Half-warp-per-environment fused step kernel for N <= 64 pedestrians, float32: TWO environments per warp.
// Same per-pedestrian arithmetic, the same compaction order and the same pairwise pass as evac_warp_kernel (evac_warp.cuh;
// reference file:line citations there), with another work mapping:
//   * lanes 0-15 run environment 2 w, lanes 16-31 environment 2 w + 1; lane l of a half owns pedestrians l, l + 16, l + 32,
//     l + 48 (N = 60: 15 of 16 lanes busy, like 60 of 64 slots before);
//   * every per-environment instruction (agent step, Philox key schedule, rewards, time, record load / store, loop control)
//     serves two environments, and in the pairwise pass one broadcast LDS.128 pair feeds FOUR targets instead of two:
//     ~16 % fewer warp-instructions per environment-step;
//   * half as many CTAs to rasterise (2048 one-warp CTAs for the 4096-environment batch) and four independent
//     pedestrians per lane of instruction-level parallelism in place of the second set of resident warps.
// Collectives are scoped to the half (member masks 0x0000ffff / 0xffff0000), so an odd environment count simply retires
// the upper half of the last warp.  The strip-culling variant of evac_warp_kernel is not reproduced here.
#pragma once
#include "evac_warp.cuh"

namespace evac {

#ifndef EVAC_HW_MINB
#define EVAC_HW_MINB 14  // resident warps per SM the register allocation is sized for: 2048 warps / 148 SMs = 13.8 (one wave)
#endif

template <typename T>
__device__ __forceinline__ T half_sum(T v, uint32_t hm) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(hm, v, o);
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(32, EVAC_HW_MINB) evac_hw_kernel(const __grid_constant__ KArgs<float> a) {  // @region hload
  constexpr int PP = 4;  // pedestrians per lane
  __shared__ __align__(16) float4 tile_all[2][66];
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int e = blockIdx.x * 2 + half, N = a.N;
  if (e >= a.E) return;  // odd batch: the upper half of the last warp leaves; every collective below names its own half only
  const uint32_t hm = 0xffffu << (16 * half);
  const Tile<float> tile(reinterpret_cast<unsigned char*>(tile_all[half]), 64);
  const uint32_t lt_mask = (1u << hl) - 1u;
  bool valid[PP];
#pragma unroll
  for (int k = 0; k < PP; ++k) valid[k] = hl + 16 * k < N;
  const uint32_t env_g = (uint32_t)(a.env_offset + e);
  const float2 exit_p = make_float2(0.f, -1.f);  // area.py:36-39

  // ---------------- load state (EnvBlock, evac_kernels.cuh): everything this launch reads is requested before anything is used
  unsigned char* const blk = a.blocks + (size_t)e * BLK_BYTES;
  float4* const ped_l = reinterpret_cast<float4*>(blk + BLK_PED) + hl;
  unsigned short* const st_l = reinterpret_cast<unsigned short*>(blk + BLK_STATUS) + hl;  // byte 2 l: pedestrian l, 2 l + 1: l + 32
  float4 pd[PP];
#pragma unroll
  for (int k = 0; k < PP; ++k) pd[k] = ped_l[16 * k];
  const unsigned st2a = st_l[0], st2b = st_l[16];
  const float4 rec0 = *reinterpret_cast<const float4*>(blk);      // agent_pos, agent_dir
  const int4 rec1 = *reinterpret_cast<const int4*>(blk + 16);     // now, episode, agent_state
  const longlong2 rec2 = *reinterpret_cast<const longlong2*>(blk + 32);
  long long overall = rec2.x;
  double acc_r = __longlong_as_double(rec2.y), acc_i = 0, acc_s = 0;
  if (hl == 0) {
    const double2 r3 = *reinterpret_cast<const double2*>(blk + 48);
    acc_i = r3.x; acc_s = r3.y;
  }
  WPed q[PP];
#pragma unroll
  for (int k = 0; k < PP; ++k) { q[k].p = make_float2(pd[k].x, pd[k].y); q[k].d = make_float2(pd[k].z, pd[k].w); }
  q[0].st = (int)(st2a & 0xffu); q[1].st = (int)(st2b & 0xffu); q[2].st = (int)(st2a >> 8); q[3].st = (int)(st2b >> 8);
  float2 ap = make_float2(rec0.x, rec0.y), ad = make_float2(rec0.z, rec0.w);
  int now = rec1.x, episode = rec1.y, wac_state = rec1.z;
  float reward_sum = 0.f;
  int any_term = 0, any_trunc = 0;
  const float noise_c = a.noise_coef;
  const float* noise_e = a.noise ? a.noise + (size_t)e * N : nullptr;
  float* obs_e = a.obs ? a.obs + (size_t)e * a.obs_dim : nullptr;
  const float2 wall = make_float2(a.width, a.height);

  for (int s = 0; s < a.num_steps; ++s) {  // @region hrng
    // ---------------- Time.step [area.py:53-59]
    const int now_prev = now;
    now += 1;
    const bool truncated = now >= a.max_timesteps;
    float2 act_tbl = make_float2(0.f, 0.f);
    if (a.agent_kind == AGENT_TABLE) act_tbl = a.actions[(size_t)s * a.E + e];
    // ---------------- angular noise [area.py:124]: pedestrian i reads word (i >> 5) & 1 of block i & 31 (philox.cuh) ->
    // blocks l and l + 16 hold the four words of this lane
    float nz[PP];
    if (noise_e != nullptr) {
      const float* np_ = noise_e + (size_t)s * a.E * N;
#pragma unroll
      for (int k = 0; k < PP; ++k) nz[k] = valid[k] ? np_[hl + 16 * k] : 0.f;
    } else {
      const uint2 r0 = evac_noise_block(a.seed, env_g, (uint32_t)episode, (uint32_t)now_prev, (uint32_t)hl);
      const uint2 r1 = evac_noise_block(a.seed, env_g, (uint32_t)episode, (uint32_t)now_prev, (uint32_t)(hl + 16));
      nz[0] = (u01(r0.x) - 0.5f) * noise_c; nz[1] = (u01(r1.x) - 0.5f) * noise_c;
      nz[2] = (u01(r0.y) - 0.5f) * noise_c; nz[3] = (u01(r1.y) - 0.5f) * noise_c;
    }
    // ---------------- escaped / exiting preparation + unit directions [area.py:79-101]  // @region hprep
    float2 u[PP];
    bool efv[PP], fv[PP];
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const int so = q[k].st;
      if (so == ST_ESCAPED) { q[k].d = make_float2(0.f, 0.f); q[k].p = exit_p; }
      float2 v = q[k].d;
      if (so == ST_EXITING) v = __fadd2_rn(exit_p, make_float2(-q[k].p.x, -q[k].p.y));
      efv[k] = (unsigned)(so - ST_VISCEK) <= (unsigned)(ST_EXITING - ST_VISCEK);
      fv[k] = (unsigned)(so - ST_VISCEK) <= (unsigned)(ST_FOLLOWER - ST_VISCEK);
      const float2 sq = __fmul2_rn(v, v);
      const float n2 = sq.x + sq.y;
      const float inv = inv_norm(n2);  // a zero direction gives NaN exactly like area.py:101
      u[k] = __fmul2_rn(v, splat(inv));
      if (so == ST_EXITING) q[k].d = __fmul2_rn(v, splat(fminf(n2 * inv, a.step_size) * inv));  // dir = u * min(|v|, step)
    }
    // ---------------- compact the moving pedestrians into the half's tile, in pedestrian-index order  // @region hcompact
    int n_src = 0;
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      const uint32_t m = (__ballot_sync(hm, efv[k]) >> (16 * half)) & 0xffffu;
      if (efv[k]) tile.put(n_src + __popc(m & lt_mask), q[k].p.x, q[k].p.y, u[k].x, u[k].y);
      n_src += __popc(m);
    }
    if (hl < 2) tile.put(n_src + hl, PARK, PARK, 0.f, 0.f);  // pad to an even count
    __syncwarp(hm);
    // ---------------- action source + Area.agent_step [area.py:182-210], IEEE float32 like the reference  // @region hagent
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
        const double ph = 0.05 * (double)(overall + 1);
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
    // ---------------- pairwise alignment [area.py:105-119]: four targets per lane against the half's own tile  // @region hpairwise
    float sx[PP], sy[PP], cnt[PP];
    {
      float xi[PP], yi[PP];
#pragma unroll
      for (int k = 0; k < PP; ++k) { xi[k] = q[k].p.x; yi[k] = q[k].p.y; }
      if (__any_sync(hm, fv[0] | fv[1] | fv[2] | fv[3])) pairwise_pass<PP, false, EVAC_PAIR_UNROLL / 2>(tile, n_src, xi, yi, a.thr2_ped, sx, sy, cnt);
      else {
#pragma unroll
        for (int k = 0; k < PP; ++k) sx[k] = sy[k] = 0.f;
      }
    }
    // ---------------- new headings, enslaving, integration, reflection, statuses  // @region hupdate
    const float2 e_ad = make_float2(__fmul_rn(a.enslaving_f, ad.x), __fmul_rn(a.enslaving_f, ad.y));  // float32 like area.py:140
    int counts = 0;  // k_exit | k_fol << 8 | n_esc << 16 | n_exi << 24   (each <= 64)
    float sum_dexit = 0.f;
#pragma unroll
    for (int k = 0; k < PP; ++k) {
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
    // ---------------- half-warp reduction, rewards, termination [reward.py:19-46, area.py:174-180, env.py:158-171]  // @region hreward
    counts = __reduce_add_sync(hm, counts);
    const float sd = half_sum(sum_dexit, hm);
    const int K_exit = counts & 0xff, K_fol = (counts >> 8) & 0xff, N_esc = (counts >> 16) & 0xff, N_exi = (counts >> 24) & 0xff;
    auto count_followers = [&]() {
      int n = 0;
#pragma unroll
      for (int k = 0; k < PP; ++k) n += __popc(__ballot_sync(hm, q[k].st == ST_FOLLOWER));
      return n;
    };
    if (a.status_counts != nullptr) {  // pedestrians.status_stats after this step (before a same-step reset)
      const int n_fol = count_followers();
      if (hl == 0)
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
    // ---------------- same-step auto-reset (gymnasium vector-env semantics)  // @region hreset
    if (a.auto_reset && (terminated || truncated)) {
      const int N_fol = count_followers();
      if (hl == 0) {  // the logging dict of env.py:115-125
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
      for (int k = 0; k < PP; ++k) {
        if (valid[k]) {
          float dummy;
          random_layout<float>(a.seed, env_g, (uint32_t)episode, (uint32_t)(hl + 16 * k), q[k].p.x, q[k].p.y, q[k].d.x, q[k].d.y);
          q[k].st = status_of<float>(q[k].p.x, q[k].p.y, 0.f, 0.f, a, dummy);
        }
      }
    }
    // ---------------- observation [wrappers.py:8-96, gravity_encoding.py:8-81]  // @region hobs
    if (obs_e != nullptr && (a.obs_every_step || s == a.num_steps - 1)) {
      float* row = obs_e + (a.obs_every_step ? (size_t)s * a.E * a.obs_dim : (size_t)0);
      if constexpr (MODE == WMODE_REL_OHE_BOX) {
        // rows = [agent; exit; pedestrians], cols = [x, y, ohe(4)]; relative positions / sqrt(2) (float32 hypotenuse)
        const float inv_hyp = (float)(1.0 / 1.41421353816986083984375);
        const float2 nap = make_float2(-ap.x, -ap.y);
        if (hl == 0) {
          const float2 ex = __fmul2_rn(__fadd2_rn(exit_p, nap), splat(inv_hyp));
          float2* r = reinterpret_cast<float2*>(row);
          r[0] = ap; r[1] = make_float2(0.f, 0.f); r[2] = make_float2(0.f, 0.f);
          r[3] = ex; r[4] = make_float2(1.f, 0.f); r[5] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < PP; ++k) {
          if (valid[k]) {
            float2* r = reinterpret_cast<float2*>(row + (hl + 16 * k + 2) * 6);
            const int st = q[k].st;
            r[0] = __fmul2_rn(__fadd2_rn(q[k].p, nap), splat(inv_hyp));
            r[1] = make_float2(st == ST_ESCAPED ? 1.f : 0.f, st == ST_EXITING ? 1.f : 0.f);
            r[2] = make_float2(st == ST_FOLLOWER ? 1.f : 0.f, st == ST_VISCEK ? 1.f : 0.f);
          }
        }
      } else if (MODE == WMODE_GRAV || a.positions == POS_GRAV) {
        // gravity encoding [gravity_encoding.py:8-38], straight-line like evac_warp_kernel; the two sums over the half in double
        float gx = 0.f, gy = 0.f;
        int nf = 0;
        const float alpha = (float)a.alpha, eps = (float)a.eps;
#pragma unroll
        for (int k = 0; k < PP; ++k) {
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
        nf = __reduce_add_sync(hm, nf);
        const double wx = half_sum((double)gx, hm), wy = half_sum((double)gy, hm);
        if (hl == 0) store_grav_obs<float>(row, ap.x, ap.y, (float)wx, (float)wy, nf, a);
      } else {
        if (hl == 0) store_head_obs<float>(row, ap.x, ap.y, a);
#pragma unroll
        for (int k = 0; k < PP; ++k)
          if (valid[k]) store_ped_obs<float>(row, hl + 16 * k, N, q[k].p.x, q[k].p.y, q[k].st, ap.x, ap.y, a);
      }
    }
    __syncwarp(hm);  // the tile is rewritten by the next step
  }  // steps

  // ---------------- write back  // @region hwriteback
#pragma unroll
  for (int k = 0; k < PP; ++k) ped_l[16 * k] = make_float4(q[k].p.x, q[k].p.y, q[k].d.x, q[k].d.y);
  st_l[0] = (unsigned short)(q[0].st | (q[2].st << 8));
  st_l[16] = (unsigned short)(q[1].st | (q[3].st << 8));
  if (a.status_out != nullptr) {  // host face: dense [E,N] statuses travel in the result block
    uint8_t* so = a.status_out + (size_t)e * N + hl;
#pragma unroll
    for (int k = 0; k < PP; ++k)
      if (valid[k]) so[16 * k] = (uint8_t)q[k].st;
  }
  if (hl == 0) {
    *reinterpret_cast<float4*>(blk) = make_float4(ap.x, ap.y, ad.x, ad.y);
    *reinterpret_cast<int4*>(blk + 16) = make_int4(now, episode, wac_state, 0);
    *reinterpret_cast<longlong2*>(blk + 32) = make_longlong2(overall, __double_as_longlong(acc_r));
    *reinterpret_cast<double2*>(blk + 48) = make_double2(acc_i, acc_s);
    if (a.reward) a.reward[e] = reward_sum;
    if (a.terminated) a.terminated[e] = (uint8_t)any_term;
    if (a.truncated) a.truncated[e] = (uint8_t)any_trunc;
  }
}

}  // namespace evac
