// mmpc_team.cuh -- the sequential part of one interior-point iteration (KKT reduction, barrier
// update, Riccati factorisation with inertia correction, roll-out of the Newton step) executed by
// a TEAM of 16 lanes per instance, two instances per warp.
//
// Column-parallel Riccati.  Lane c of the team owns one column of the stage system over (x, u):
//     c = 0..8   x-columns      c = 9..13   u-columns      c = 14   right-hand side      c = 15 idle
// The cost-to-go P(k+1) lives in the registers of lanes 0..8 (one column each) for the whole backward
// sweep.  A stage is:  Y = P [A B]  (the few columns A and B couple are fetched from the neighbour
// lanes with width-16 shuffles),  M = [A B]^T Y + H  (local),  rank-one elimination of the stage slack
// v_k = s_{k+1}  (local),  right-looking LDL^T of the u block (the pivot column is broadcast, every
// lane updates its own column, so P(k) and p(k) simply remain in lanes 0..8 and 14), back-substitution
// for the gains (local).  Per stage and lane: ~200 DFMA, ~190 double shuffles, no shared memory, no
// local memory, ~20x shorter dependency chain than the one-thread-per-instance recursion -- this
// kernel sets the latency of a round once the batch has thinned out.
//
// Every decision (pivot signs, convergence, barrier update) is taken on values that are bit-identical
// in all 16 lanes (broadcast values, butterfly reductions), so a team never diverges at a shuffle, and
// the two teams of a warp run the same instruction stream (full-mask, width-16 shuffles; a team that
// needs no further factorisation repeats its last one while its partner retries with a larger delta_w).
#pragma once
#include "mmpc_staged.cuh"

namespace mmpc {

// index into the packed stage-QP record of entry (r, c) of the (x, u) Hessian, -1 if structurally zero
__host__ __device__ constexpr int team_hidx(int r, int c) {
  if (r > c) { int t = r; r = c; c = t; }
  const int X2P[9] = {0, 1, 2, -1, -1, -1, 3, 4, 5};
  if (c < 9) {
    int pr = X2P[r], pc = X2P[c];
    if (pr >= 0 && pc >= 0) return Q_HP + pidx(pr, pc);
    if (pr < 0 && pc < 0) {
      if (r == c) return Q_HVD + (r - 3);
      if (r == 3 && c == 5) return Q_H35;
      if (r == 4 && c == 5) return Q_H45;
    }
    return -1;
  }
  if (r >= 9) return r == c ? Q_HUU + (r - 9) : -1;
  return (r == 2 && c == 9) ? Q_HPU : -1;
}
// Slot of the stage-QP record (in the shared-memory ring) that lane c adds to row r of its column: the Hessian entry for
// the x and u lanes, the gradient entry for the right-hand-side lane, BW_ZERO (a slot that holds 0.0) where the entry is
// structurally zero.  Packed one byte per row into four words per lane, which the lane keeps in registers: the hot loop
// then is  M[r] += q[slot(r)]  without a table look-up, a branch on the sign of the index, or a diverging right-hand side.
constexpr int TEAM_BW_ZERO = 95;
__host__ __device__ constexpr int team_slot(int r, int c) {
  if (c < 14) { const int i = team_hidx(r, c); return i >= 0 ? i : TEAM_BW_ZERO; }
  if (c == 14) return Q_GA + (r < 9 ? r : r + 1);
  return TEAM_BW_ZERO;
}
struct TeamPack {
  unsigned v[16][4];
  constexpr TeamPack() : v() {
    for (int c = 0; c < 16; ++c)
      for (int r = 0; r < 14; ++r) v[c][r >> 2] |= (unsigned)team_slot(r, c) << (8 * (r & 3));
  }
};
__device__ constexpr TeamPack TEAM_PACK = TeamPack();

struct Team {
  Inst S;
  unsigned hp[4];         // TEAM_PACK.v[c]: the record slots of this lane's column
  int c;                  // lane of the team
  int sd;                 // slot of the diagonal entry of this lane's column (x and u lanes)
  __device__ __forceinline__ Team(const SParams& p, int b, int c_, double* sm_) : S(p, b), c(c_), sm(sm_) {
#pragma unroll
    for (int w = 0; w < 4; ++w) hp[w] = TEAM_PACK.v[c_][w];
    sd = c_ < 14 ? (int)((TEAM_PACK.v[c_][c_ >> 2] >> (8 * (c_ & 3))) & 0xffu) : TEAM_BW_ZERO;
  }
  __device__ __forceinline__ int slot(int r) const { return (int)((hp[r >> 2] >> (8 * (r & 3))) & 0xffu); }  // r: compile-time

  __device__ __forceinline__ static double tsum(double v) {
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) v += shfl16_xor(v, m);
    return v;
  }
  __device__ __forceinline__ static double tmax(double v) {
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) v = fmax(v, shfl16_xor(v, m));
    return v;
  }
  __device__ __forceinline__ static double tmin(double v) {
#pragma unroll
    for (int m = 8; m > 0; m >>= 1) v = fmin(v, shfl16_xor(v, m));
    return v;
  }

  // m[0..13] = [A B]^T y
  __device__ __forceinline__ static void abt_mul(const double (&y)[9], const SACoef& a, double (&m)[14]) {
    m[0] = y[0]; m[1] = y[1];
    m[2] = y[2] + a.a32 * y[3] + a.a42 * y[4];
    m[3] = y[3] + a.dt * y[0] + a.a43 * y[4];
    m[4] = y[4] + a.dt * y[1] + a.a34 * y[3];
    m[5] = y[5] + a.dt * y[2] + a.a35 * y[3] + a.a45 * y[4];
    m[6] = y[6]; m[7] = y[7]; m[8] = y[8];
    m[9] = a.dt * (a.cp * y[3] + a.sp * y[4]); m[10] = a.dt * y[5];
    m[11] = a.dt * y[6]; m[12] = a.dt * y[7]; m[13] = a.dt * y[8];
  }

  // ---- shared-memory prefetch ring (cp.async) ---------------------------------------------------------
  // backward sweep, one record per stage:  QP[80] | dfc[9] | u0 x3 x4 x5 cos sin | pad   (2 slots)
  // roll-out, one record per stage:        RK[120] | a[6] | dfc[9] | u0 x3 x4 x5 cos sin | pad   (4 slots)
  static constexpr int BW_DFC = 80, BW_AC = 89, BW_SZ = 96, BW_SLOTS = 2;  // [95] = TEAM_BW_ZERO, never written by the copies
  static constexpr int RO_A = 120, RO_DFC = 126, RO_AC = 135, RO_SZ = 144, RO_SLOTS = 4;
  static constexpr int SMEM_DOUBLES = RO_SZ * RO_SLOTS;  // per team (>= BW_SZ * BW_SLOTS)
  double* sm;  // this team's ring

  // the dynamics coefficients of stage k: lanes 9..14 fetch u0 x3 x4 x5 cos sin
  __device__ __forceinline__ void fetch_ac(double* dst, int k, int it) {
    if (c >= 9 && c < 15) {
      const int q = c - 9;
      const double* src = q == 0 ? &S.W(k, it + I_U + 0) : q < 4 ? &S.W(k, it + I_X + 2 + q) : &S.W2(k, S_FK + (q - 4));
      async_copy8(dst + q, src);
    }
  }
  __device__ __forceinline__ static SACoef ac_from(const double* q, double dt) {
    SACoef a; a.dt = dt; a.cp = q[4]; a.sp = q[5];
    const double u0 = q[0], x3 = q[1], x4 = q[2], x5 = q[3];
    a.a32 = -dt * u0 * a.sp; a.a42 = dt * u0 * a.cp; a.a34 = -dt * x5; a.a43 = dt * x5; a.a35 = -dt * x4; a.a45 = dt * x3;
    return a;
  }
  __device__ __forceinline__ void bw_issue(int k, int it) {
    if (k >= 0) {
      double* dst = sm + (k & 1) * BW_SZ;
      const double* src = &S.Qw(k, 0);
#pragma unroll
      for (int r = 0; r < 3; ++r) { int i = c + 16 * r; if (i < QS / 2) async_copy16(dst + 2 * i, src + 2 * i); }
      if (k < S.N) {
        if (c < 9) async_copy8(dst + BW_DFC + c, &S.W2(k, S_DFC + c));
        fetch_ac(dst + BW_AC, k, it);
      }
    }
    async_commit();
  }
  __device__ __forceinline__ void ro_issue(int k, int it) {
    if (k <= S.N) {
      double* dst = sm + (k & 3) * RO_SZ;
      const double* src = &S.Rw(k, 0);
#pragma unroll
      for (int r = 0; r < 4; ++r) { int i = c + 16 * r; if (i < RS / 2) async_copy16(dst + 2 * i, src + 2 * i); }
      if (c < 3) async_copy16(dst + RO_A + 2 * c, &S.Qw(k, Q_A + 2 * c));
      if (k < S.N) {
        if (c < 9) async_copy8(dst + RO_DFC + c, &S.W2(k, S_DFC + c));
        fetch_ac(dst + RO_AC, k, it);
      }
    }
    async_commit();
  }

  // Backward sweep.  Returns 0, or 1 on a non-positive pivot (wrong inertia).
  template <bool Q3>
  __device__ int riccati(double reg, double mu, int it) {
    const int N = S.N;
    const bool rhs = (c == 14);
    double Pc[9];            // lanes 0..8: column c of Pxx(k+1); lane 14: pxx(k+1)
    double an[9], cn, gsn;   // slack column of stage k+1 (replicated): H[x][s] (pose entries, except after the terminal-row stage), H[s][s], g_s
    // MMPC_MODE_REFERENCE to the letter (Inst::q3): the terminal self-collision rows couple x_N with s[N-1].  Their column
    // (abar, cbar, gbar; v-slots of stage N's record) enters stage N-1 as [A B]^T abar, i.e. s[N-1] is coupled with u[N-1]
    // and has to take part in that stage's control elimination: lane 15 carries the slack column through the LDL^T
    // like the x lanes carry theirs, `ss` is the s-row of lanes 15 (H[s][s]) and 14 (g_s).
    constexpr bool q3 = Q3;  // compile-time: the clean / rows-on-s[N] kernels carry none of this
    double ab[NP], cb = 0, gb = 0;
    team_sync();             // the ring may still be read by a slower lane of the previous phase
    if (c == 0) { sm[TEAM_BW_ZERO] = 0.0; sm[BW_SZ + TEAM_BW_ZERO] = 0.0; }  // (the roll-out reuses the ring with another layout)
    bw_issue(N, it); bw_issue(N - 1, it);
    async_wait<1>(); team_sync();
    {
      double* q = sm + (N & 1) * BW_SZ;
      if (c < 9) q[sd] += reg;  // the diagonal entry of a column is read by its own lane only
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int i1 = slot(r), i2 = rhs ? i1 + (Q_GB - Q_GA) : TEAM_BW_ZERO;
        const double v = fma(mu, q[i2], q[i1]);
        Pc[r] = (c < 9 || rhs) ? v : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) an[i] = 0;
#pragma unroll
      for (int a = 0; a < NP; ++a) { an[POSE2X[a]] = q[Q_A + a]; ab[a] = q3 ? q[Q_BV + a] : 0.0; }
      cn = q[Q_C]; gsn = q[Q_GA + SGY_S] + mu * q[Q_GB + SGY_S];
      if (q3) { cb = q[Q_HVV]; gb = q[Q_GA + SGY_V] + mu * q[Q_GB + SGY_V]; }
      if (c < 9) {
#pragma unroll
        for (int r = 0; r < 9; ++r) if (r <= c) S.Rw(N, R_P + ssidx(r, c)) = Pc[r];
      } else if (rhs) {
#pragma unroll
        for (int r = 0; r < 9; ++r) S.Rw(N, R_PV + r) = Pc[r];
      }
    }
    team_sync();
    bw_issue(N - 2, it);
    int bad = 0;
    // which neighbour columns this lane's column of P [A B] needs (fixed per lane; only the coefficients change)
    int s1 = c, s2 = c, s3 = c;
    switch (c) {
      case 2: s1 = 3; s2 = 4; break;
      case 3: s1 = 0; s2 = 4; break;
      case 4: s1 = 1; s2 = 3; break;
      case 5: s1 = 2; s2 = 3; s3 = 4; break;
      case 9: s1 = 3; s2 = 4; break;
      case 10: s1 = 5; break;
      case 11: s1 = 6; break;
      case 12: s1 = 7; break;
      case 13: s1 = 8; break;
      default: break;
    }
    for (int k = N - 1; k >= 0; --k) {
      async_wait<1>(); team_sync();
      double* qm = sm + (k & 1) * BW_SZ;
      const double* q = qm;
      double* rkk = &S.Rw(k, 0);
      const SACoef a = ac_from(q + BW_AC, S.dt);
      double d[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) d[i] = q[BW_DFC + i];
      // (Pxx d)[c] in lane c (symmetry), gathered into the right-hand-side lane: pd = pxx + Pxx d
      double zd = 0;
#pragma unroll
      for (int i = 0; i < 9; ++i) zd = fma(Pc[i], d[i], zd);
      double Y[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) { double g = shfl16(zd, i); Y[i] = Pc[i] + (rhs ? g : 0.0); }  // u lanes: Pc = 0
      // Y += coef * Pxx[:, src]: the columns A and B couple (A = I + sparse, B sparse)
      // coefficients of those columns: A = I + sparse, B sparse
      const bool dtc = (c >= 3 && c <= 5) || (c >= 10 && c <= 13);
      const double c1 = c == 2 ? a.a32 : c == 9 ? a.dt * a.cp : dtc ? a.dt : 0.0;
      const double c2 = c == 2 ? a.a42 : c == 3 ? a.a43 : c == 4 ? a.a34 : c == 5 ? a.a35 : c == 9 ? a.dt * a.sp : 0.0;
      const double c3 = c == 5 ? a.a45 : 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        double g1 = shfl16(Pc[i], s1), g2 = shfl16(Pc[i], s2), g3 = shfl16(Pc[i], s3);
        Y[i] = fma(c1, g1, fma(c2, g2, fma(c3, g3, Y[i])));
      }
      double M[14];
      abt_mul(Y, a, M);
      // + H(:, c) + reg on the diagonal; right-hand side: + g_A + mu g_B -- one code path for all lanes (slot(), BW_ZERO)
      if (c < 14) qm[sd] += reg;  // the diagonal entry of a column is read by its own lane only
#pragma unroll
      for (int r = 0; r < 14; ++r) {
        const int i1 = slot(r), i2 = rhs ? i1 + (Q_GB - Q_GA) : TEAM_BW_ZERO;
        M[r] += fma(mu, q[i2], q[i1]);
      }
      // eliminate v_k = s_{k+1}:  w = [A B]^T a(k+1) + bv(k),  cv = hvv(k) + c(k+1),  l0 = g_s(k+1) + a.d + g_v(k)
      double wv[14];
      abt_mul(an, a, wv);
#pragma unroll
      for (int p = 0; p < NP; ++p) wv[POSE2X[p]] += q[Q_BV + p];
      double l0 = gsn + q[Q_GA + SGY_V] + mu * q[Q_GB + SGY_V];
#pragma unroll
      for (int i = 0; i < 9; ++i) l0 = fma(an[i], d[i], l0);
      const double cv = q[Q_HVV] + cn, icv = rcp(cv);
      bad |= !(cv > 1e-13);
      double wc = 0;
#pragma unroll
      for (int r = 0; r < 14; ++r) wc = (r == c) ? wv[r] : wc;
      {
        const double fac = (rhs ? l0 : wc) * icv;
#pragma unroll
        for (int r = 0; r < 14; ++r) M[r] = fma(-wv[r], fac, M[r]);
      }
      // terminal-row stage (q3, k = N-1): lane 15 takes the slack column  H[(x,u)][s] = a(k) + [A B]^T abar,  ss = its s-row
      const bool q3k = q3 && k == N - 1;
      double ss = 0;
      if (q3k) {
        double abe[9], wb[14];
#pragma unroll
        for (int i = 0; i < 9; ++i) abe[i] = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) abe[POSE2X[p]] = ab[p];
        abt_mul(abe, a, wb);
        double abd = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) { wb[POSE2X[p]] += q[Q_A + p]; abd = fma(ab[p], d[POSE2X[p]], abd); }
        if (c == 15) {
#pragma unroll
          for (int r = 0; r < 14; ++r) M[r] = wb[r];
          ss = q[Q_C] + cb;
        } else if (rhs) ss = q[Q_GA + SGY_S] + mu * q[Q_GB + SGY_S] + gb + abd;
      }
      // this stage's slack column is the next iteration's a(k+1); the ring slot is then free
#pragma unroll
      for (int i = 0; i < 9; ++i) an[i] = 0;
#pragma unroll
      for (int p = 0; p < NP; ++p) an[POSE2X[p]] = q[Q_A + p];
      cn = q[Q_C]; gsn = q[Q_GA + SGY_S] + mu * q[Q_GB + SGY_S];
      team_sync();
      bw_issue(k - 2, it);
      // right-looking LDL^T of the u block: broadcast the pivot column, update the own column
      double yv[NU], Lm[NU][NU];
#pragma unroll
      for (int p = 0; p < NU; ++p) {
        const int pl = 9 + p;
        double col[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) if (r < 9 || r >= pl) col[r] = shfl16(M[r], pl);
        const double piv = col[pl];
        bad |= !(piv > 1e-13);
        const double ip = rcp(piv);
        const double f = M[pl] * ip;
        yv[p] = f;
        if (q3k) ss = fma(-shfl16(M[pl], 15), f, ss);  // s-row of this lane's column (by symmetry the pivot column's s entry is lane 15's)
#pragma unroll
        for (int r = 0; r < 14; ++r) if (r < 9 || r > pl) M[r] = fma(-col[r], f, M[r]);
#pragma unroll
        for (int q2 = p + 1; q2 < NU; ++q2) Lm[q2][p] = col[9 + q2] * ip;
      }
      if (warp_all(bad != 0)) { async_wait<0>(); return 1; }  // both teams of the warp failed: no point in going on
      // gains: K(:, c) = -L^-T y  (lane 14: kff)
      double kc[NU];
#pragma unroll
      for (int p = NU - 1; p >= 0; --p) {
        double v = yv[p];
#pragma unroll
        for (int q2 = p + 1; q2 < NU; ++q2) v = fma(-Lm[q2][p], kc[q2], v);
        kc[p] = v;
      }
      if (c < 9) {
#pragma unroll
        for (int p = 0; p < NU; ++p) rkk[R_K + p * NX + c] = -kc[p];
#pragma unroll
        for (int r = 0; r < 9; ++r) if (r <= c) rkk[R_P + ssidx(r, c)] = M[r];
      } else if (rhs) {
#pragma unroll
        for (int p = 0; p < NU; ++p) rkk[R_KFF + p] = -kc[p];
#pragma unroll
        for (int r = 0; r < 9; ++r) rkk[R_PV + r] = M[r];
        rkk[R_CV] = cv; rkk[R_L0] = l0;
      }
      if (c < 14) rkk[R_W + c] = wc;
      if (q3k) {  // the slack column after the control elimination: what stage N-2 eliminates v = s[N-1] with
#pragma unroll
        for (int r = 0; r < 9; ++r) an[r] = shfl16(M[r], 15);
        cn = shfl16(ss, 15); gsn = shfl16(ss, 14);
        bad |= !(cn > 1e-13);
        if (c == 15) {  // gain of u[N-1] on ds[N-1] and the column itself, for the roll-out (free slots of stage N's record)
          double* rkn = &S.Rw(N, 0);
#pragma unroll
          for (int p = 0; p < NU; ++p) rkn[R_K + p] = -kc[p];
#pragma unroll
          for (int r = 0; r < 9; ++r) rkn[R_K + 8 + r] = M[r];
        }
      }
#pragma unroll
      for (int r = 0; r < 9; ++r) Pc[r] = (c < 9 || rhs) ? M[r] : 0.0;
    }
    async_wait<0>();
    if (!(cn > 1e-13)) return 1;
    return bad;
  }

  // roll-out of the Newton step: dx, du, ds and the new costates lam+ = P [dx; ds] + p.
  // dx is replicated in the team; lane a < 5 forms du[a], lane i < 9 forms lam+[i].
  template <bool Q3>
  __device__ void rollout(double mu, int it) {
    const int N = S.N; const double dt = S.dt;
    team_sync();
    ro_issue(0, it); ro_issue(1, it); ro_issue(2, it); ro_issue(3, it);
    double dxv[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) dxv[i] = 0;
    if (c < 9) S.W2(0, S_DX + c) = 0;
    double dsv = -(S.Qw(0, Q_GA + SGY_S) + mu * S.Qw(0, Q_GB + SGY_S)) / S.Qw(0, Q_C);
    if (c == 14) S.W2(0, S_DS) = dsv;
    constexpr bool q3 = Q3;
    for (int k = 0; k < N; ++k) {
      async_wait<2>(); team_sync();
      const double* r0 = sm + (k & 3) * RO_SZ;
      const double* r1 = sm + ((k + 1) & 3) * RO_SZ;
      double* w2k = S.stage_ptr(k, S.B2); double* w2n = S.stage_ptr(k + 1, S.B2);
      const double ds_in = dsv;  // ds[k]
      double mine = 0;
      if (c < NU) {
        mine = r0[R_KFF + c];
#pragma unroll
        for (int j = 0; j < NX; ++j) mine = fma(r0[R_K + c * NX + j], dxv[j], mine);
        if (q3 && k == N - 1) mine = fma(S.Rw(N, R_K + c), ds_in, mine);  // u[N-1] also answers to ds[N-1] (terminal rows)
        w2k[(S_DU + c) << LSH] = mine;
      }
      double duv[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) duv[a] = shfl16(mine, a);
      double l = r0[R_L0];
#pragma unroll
      for (int i = 0; i < NX; ++i) l = fma(r0[R_W + i], dxv[i], l);
#pragma unroll
      for (int a = 0; a < NU; ++a) l = fma(r0[R_W + NX + a], duv[a], l);
      dsv = -l * rcp(r0[R_CV]);
      const SACoef a = ac_from(r0 + RO_AC, dt);
      double nx_[NX];
      nx_[0] = dxv[0] + dt * dxv[3]; nx_[1] = dxv[1] + dt * dxv[4]; nx_[2] = dxv[2] + dt * dxv[5];
      nx_[3] = dxv[3] + a.a32 * dxv[2] + a.a34 * dxv[4] + a.a35 * dxv[5] + dt * a.cp * duv[0];
      nx_[4] = dxv[4] + a.a42 * dxv[2] + a.a43 * dxv[3] + a.a45 * dxv[5] + dt * a.sp * duv[0];
      nx_[5] = dxv[5] + dt * duv[1];
      nx_[6] = dxv[6] + dt * duv[2]; nx_[7] = dxv[7] + dt * duv[3]; nx_[8] = dxv[8] + dt * duv[4];
#pragma unroll
      for (int i = 0; i < NX; ++i) dxv[i] = nx_[i] + r0[RO_DFC + i];
      if (c < 9) {
        double mx = 0;
#pragma unroll
        for (int i = 0; i < NX; ++i) mx = (i == c) ? dxv[i] : mx;
        w2n[(S_DX + c) << LSH] = mx;
        double v = r1[R_PV + c];
#pragma unroll
        for (int j = 0; j < NX; ++j) v = fma(r1[R_P + (c <= j ? c * 9 - c * (c - 1) / 2 + (j - c) : j * 9 - j * (j - 1) / 2 + (c - j))], dxv[j], v);
        if (q3 && k + 1 == N - 1) v = fma(S.Rw(N, R_K + 8 + c), dsv, v);  // the column left by stage N-1's control elimination
        else {
          if (c < 3) v = fma(r1[RO_A + c], dsv, v);
          if (c >= 6) v = fma(r1[RO_A + (c - 3)], dsv, v);
        }
        if (q3 && k + 1 == N && (c < 3 || c >= 6)) v = fma(S.Qw(N, Q_BV + (c < 3 ? c : c - 3)), ds_in, v);  // abar ds[N-1]
        w2n[(S_LAMN + c) << LSH] = v;
      }
      if (c == 14) w2n[S_DS << LSH] = dsv;
      team_sync();  // slot k & 3 is free
      ro_issue(k + 4, it);
    }
    async_wait<0>();
    if (c < 2 && S.term_eq(N)) {  // Newton target of the terminal-equality multipliers
      double cq = S.W(N, it + I_X + c) - S.W2(N, IN_XREF + c);
      double dxc = c == 0 ? dxv[0] : dxv[1];
      S.W2(0, S_LAMN + c) = S.W(0, it + I_LAM + c) + (dxc + cq) / MMPC_DELTA_C;
    }
  }

  // KKT reduction, convergence test, barrier update, factorisation with inertia correction, roll-out.
  // Both teams of a warp execute exactly the same instruction stream (full-mask shuffles): a team whose
  // instance needs no (further) factorisation simply repeats its last one, which rewrites the same values.
  // An instance that leaves here is only marked (state ST_FINISH + status); its results are written by
  // the stage-parallel step kernel of the same round.
  template <bool Q3>
  __device__ void solve() {
    const MmpcConfig& cfg = S.cfg;
    const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5;
    const double tol = cfg.tol;
    const int N = S.N;
    const int it = S.J(J_CUR) * S.ITSZ;
    KktParts kp;
    kp.e_stat = 0; kp.e_prim = 0; kp.c_hi = -1e300; kp.c_lo = 1e300; kp.sum_lam = 0; kp.sum_z = 0;
    double nz = 0, neq = 0;
    for (int k = c; k <= N; k += 16) {
      kp.e_stat = fmax(kp.e_stat, S.W2(k, S_PART + 0)); kp.e_prim = fmax(kp.e_prim, S.W2(k, S_PART + 1));
      kp.c_hi = fmax(kp.c_hi, S.W2(k, S_PART + 2)); kp.c_lo = fmin(kp.c_lo, S.W2(k, S_PART + 3));
      kp.sum_lam += S.W2(k, S_PART + 4); kp.sum_z += S.W2(k, S_PART + 5); nz += S.W2(k, S_PART + 6); neq += S.W2(k, S_PART + 7);
    }
    kp.e_stat = tmax(kp.e_stat); kp.e_prim = tmax(kp.e_prim); kp.c_hi = tmax(kp.c_hi); kp.c_lo = tmin(kp.c_lo);
    kp.sum_lam = tsum(kp.sum_lam); kp.sum_z = tsum(kp.sum_z); nz = tsum(nz); neq = tsum(neq);
    kp.n_z = (int)nz; kp.n_eq = (int)neq;
    if (Q3) kp.e_stat = fmax(kp.e_stat, fabs(S.W2(N, S_DFC + 1) - S.W2(N, S_DFC + 0)));  // d/ds[N-1]: two threads' shares
    const double E0 = kkt_error(kp, 0.0);
    // every lane reads the scalar state before lane 0 rewrites any of it
    const int iter = S.J(J_IT);
    double mu = S.D(D_MU);
    const double reg_last = S.D(D_REGLAST);
    team_sync();
    if (c == 0) S.D(D_E0) = E0;
    int fin = -1;  // status with which the instance leaves the solve, -1: goes on
    if (!(E0 == E0)) fin = MMPC_STATUS_NAN;
    else if (E0 <= tol) fin = MMPC_STATUS_CONVERGED;
    else if (iter >= cfg.max_iter) fin = MMPC_STATUS_MAX_ITER;
    if (fin < 0) {
      bool mu_changed = false;
      while (kkt_error(kp, mu) <= kap_eps * mu && mu > tol / 10) {
        mu = fmax(tol / 10, fmin(kap_mu * mu, pow(mu, th_mu))); mu_changed = true;
      }
      if (mu_changed && c == 0) { S.J(J_NFILT) = 0; S.J(J_FRST) = S.J(J_FRST) & 0xff00; S.D(D_MU) = mu; }
    }
    bool need = fin < 0;
    double reg = 0;
    int tries = 0;
    for (;;) {
      const int fail = riccati<Q3>(reg, mu, it);
      if (need) {
        if (!fail) { need = false; if (c == 0) { if (reg > 0) S.D(D_REGLAST) = reg; S.J(J_REGF) = reg > 0; } }
        else {
          if (reg == 0) reg = (reg_last == 0) ? 1e-4 : fmax(1e-20, reg_last / 3);
          else reg *= (reg_last == 0 ? 100 : 8);
          if (++tries > 40 || reg > 1e20) { fin = MMPC_STATUS_FACTOR; need = false; }
        }
      }
      if (!warp_any(need)) break;
    }
    team_sync();  // the Riccati records written by the other lanes are read back in the roll-out
    rollout<Q3>(mu, it);
    if (fin >= 0 && c == 0) { S.J(J_STATUS) = fin; S.J(J_STATE) = ST_FINISH; }
  }

#ifdef MMPC_RESIDENT
  // Resident build (mmpc_resident.cu): the warp works on ONE instance, so the second half-warp has nothing of its own to do.
  // Instead of mirroring the first it factorises, at the same time and into its own set of Riccati records (rk1), with the
  // NEXT delta_w of IPOPT's inertia-correction sequence.  The sequence is known up front (0, then reg_last/3 or 1e-4, then
  // x8 or x100), the first success in sequence order wins, so the accepted factorisation -- and every number after it -- is
  // the one the serial retry loop of solve() arrives at; 26 % of the iterations of config 3 need delta_w > 0 and save one
  // full sweep or more.  rk1 lies inside the copy of the iterate that is dead between an accepted trial and the next one
  // (stage stride rk1_stride), so the second set costs no shared memory.  An instance that leaves (converged, max_iter, NaN) skips factorisation and roll-out altogether:
  // nothing reads the step of a finished instance.
  template <bool Q3>
  __device__ void solve_spec(int half, double* rk1, int rk1_stride) {
    const MmpcConfig& cfg = S.cfg;
    const double kap_eps = 10, kap_mu = 0.2, th_mu = 1.5;
    const double tol = cfg.tol;
    const int N = S.N;
    const int it = S.J(J_CUR) * S.ITSZ;
    KktParts kp;
    kp.e_stat = 0; kp.e_prim = 0; kp.c_hi = -1e300; kp.c_lo = 1e300; kp.sum_lam = 0; kp.sum_z = 0;
    double nz = 0, neq = 0;
    for (int k = c; k <= N; k += 16) {
      kp.e_stat = fmax(kp.e_stat, S.W2(k, S_PART + 0)); kp.e_prim = fmax(kp.e_prim, S.W2(k, S_PART + 1));
      kp.c_hi = fmax(kp.c_hi, S.W2(k, S_PART + 2)); kp.c_lo = fmin(kp.c_lo, S.W2(k, S_PART + 3));
      kp.sum_lam += S.W2(k, S_PART + 4); kp.sum_z += S.W2(k, S_PART + 5); nz += S.W2(k, S_PART + 6); neq += S.W2(k, S_PART + 7);
    }
    kp.e_stat = tmax(kp.e_stat); kp.e_prim = tmax(kp.e_prim); kp.c_hi = tmax(kp.c_hi); kp.c_lo = tmin(kp.c_lo);
    kp.sum_lam = tsum(kp.sum_lam); kp.sum_z = tsum(kp.sum_z); nz = tsum(nz); neq = tsum(neq);
    kp.n_z = (int)nz; kp.n_eq = (int)neq;
    if (Q3) kp.e_stat = fmax(kp.e_stat, fabs(S.W2(N, S_DFC + 1) - S.W2(N, S_DFC + 0)));
    const double E0 = kkt_error(kp, 0.0);
    const int iter = S.J(J_IT);
    double mu = S.D(D_MU);
    const double reg_last = S.D(D_REGLAST);
    team_sync();
    const bool writer = c == 0 && half == 0;
    if (writer) S.D(D_E0) = E0;
    int fin = -1;
    if (!(E0 == E0)) fin = MMPC_STATUS_NAN;
    else if (E0 <= tol) fin = MMPC_STATUS_CONVERGED;
    else if (iter >= cfg.max_iter) fin = MMPC_STATUS_MAX_ITER;
    // (every branch that guards shuffles is taken on a VOTE: the values are the same in all lanes, but only a vote result is
    // warp-uniform to the compiler, which otherwise brackets each shuffle with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE)
    if (warp_all(fin >= 0)) { if (writer) { S.J(J_STATUS) = fin; S.J(J_STATE) = ST_FINISH; } return; }
    bool mu_changed = false;
    while (kkt_error(kp, mu) <= kap_eps * mu && mu > tol / 10) {
      mu = fmax(tol / 10, fmin(kap_mu * mu, pow(mu, th_mu))); mu_changed = true;
    }
    if (mu_changed && writer) { S.J(J_NFILT) = 0; S.J(J_FRST) = S.J(J_FRST) & 0xff00; S.D(D_MU) = mu; }
    auto next_reg = [&](double r) { return r == 0 ? ((reg_last == 0) ? 1e-4 : fmax(1e-20, reg_last / 3)) : r * (reg_last == 0 ? 100 : 8); };
    double* const rk0 = S.rkp; double* const ring0 = sm;
    S.rkp = half ? rk1 : rk0; S.rks = half ? rk1_stride : RS;
    sm = ring0 + half * (BW_SZ * BW_SLOTS);   // backward sweeps: a ring per half (each adds its own delta_w to the diagonals)
    double r_a = 0, reg = 0;   // attempt number ia of the serial sequence runs in half 0, ia + 1 in half 1
    int ia = 0, win = -1;
    for (;;) {
      const double r_b = next_reg(r_a);
      const int fail = riccati<Q3>(half ? r_b : r_a, mu, it);
      const int fail_a = __shfl_sync(FULL, fail, 0), fail_b = __shfl_sync(FULL, fail, 16);
      if (warp_all(!fail_a)) { win = 0; reg = r_a; break; }
      if (warp_all(ia + 1 > 40 || r_b > 1e20)) break;      // the serial loop gives up after attempt ia (MMPC_STATUS_FACTOR)
      if (warp_all(!fail_b)) { win = 1; reg = r_b; break; }
      r_a = next_reg(r_b); ia += 2;
      if (warp_all(ia > 40 || r_a > 1e20)) break;
    }
    sm = ring0;                                // roll-out: both halves do the same, one ring
    if (warp_all(win < 0)) { if (writer) { S.J(J_STATUS) = MMPC_STATUS_FACTOR; S.J(J_STATE) = ST_FINISH; } S.rkp = rk0; S.rks = RS; return; }
    if (writer) { if (reg > 0) S.D(D_REGLAST) = reg; S.J(J_REGF) = reg > 0; }
    S.rkp = win ? rk1 : rk0; S.rks = win ? rk1_stride : RS;
    team_sync();
    rollout<Q3>(mu, it);
    S.rkp = rk0; S.rks = RS;
  }
#endif
};

// Q3: the literal reference NLP (Inst::q3) -- the host picks the instantiation from the configuration
template <bool Q3>
__device__ inline void body_solve_team(const SParams& P, int j, int c, double* sm) {
  Team T(P, list_S(P)[j], c, sm);
  T.template solve<Q3>();
}

#if !defined(MMPC_EMULATE_LANE) && !defined(MMPC_RESIDENT)
// threads per block of the team kernel, 16 lanes per instance.  A/B on B200 (solve phase of a 65,536 batch): 128 -> 162.5 ms,
// 64 -> 155.5 ms, 32 -> 153.3 ms: a block leaves when its slowest warp (most delta_w retries) does, smaller blocks free their
// registers sooner
#ifndef MMPC_TEAM_BLOCK
#define MMPC_TEAM_BLOCK 32
#endif
// WARPS = resident warps per SM the instantiation is compiled for, i.e. its register budget.  A/B on B200 (solve phase of a
// 65,536 batch / one instance end to end): 16 warps (128 registers, spills) 142.7 ms / 4.76 ms, 12 warps (168) 134.3 / 4.74,
// 10 warps 144.0 / 4.66, 8 warps (255 registers, no spills) 147.4 / 4.40: the bulk wants 12, a round whose instances all fit
// on the GPU at 8 warps per SM is pure latency and wants the spill-free code.
#ifndef MMPC_TEAM_WARPS_BULK
#define MMPC_TEAM_WARPS_BULK 12
#endif
constexpr int TEAM_WARPS_BULK = MMPC_TEAM_WARPS_BULK, TEAM_WARPS_THIN = 8;
template <bool Q3, int WARPS>
__global__ void __launch_bounds__(MMPC_TEAM_BLOCK, WARPS * 32 / MMPC_TEAM_BLOCK) staged_solve_team_kernel(const __grid_constant__ SParams P) {
  __shared__ __align__(16) double ring[(MMPC_TEAM_BLOCK / 16) * Team::SMEM_DOUBLES];  // one ring per team
  const int n = P.cnt[0];
  const int c = threadIdx.x & 15;
  // a warp strides over pairs of list entries; with an odd count the last warp's second team repeats the
  // first team's instance (same computation, same stores) so that the warp stays in lock step
  // (the warp index is formed from blockIdx alone when a block is one warp: a loop whose start depends on threadIdx is
  // "divergent" to the compiler, which then brackets every shuffle of the body with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE --
  // 296 brackets, two more instructions per double shuffle)
  const int warps = gridDim.x * (MMPC_TEAM_BLOCK / 32);
  const int warp0 = blockIdx.x * (MMPC_TEAM_BLOCK / 32) + (MMPC_TEAM_BLOCK > 32 ? (int)(threadIdx.x >> 5) : 0);
  for (int j0 = warp0 * 2; j0 < n; j0 += 2 * warps) {
    const int j = min(j0 + ((threadIdx.x >> 4) & 1), n - 1);
    body_solve_team<Q3>(P, j, c, ring + (threadIdx.x >> 4) * Team::SMEM_DOUBLES);
  }
}
#endif

}  // namespace mmpc
