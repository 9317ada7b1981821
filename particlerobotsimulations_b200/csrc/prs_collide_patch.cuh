/*
 * prs_collide_patch.cuh — collide with every pair of two patch robots evaluated ONCE (north_star (3): the
 * cell tile and its neighbouring cells staged in shared memory by TMA bulk copies).  Included by
 * prs_kernels.cu after prs_collide.cuh.
 *
 * Result reproduced: the same as k_collide_exact (collideD + collideCell + collideSpheres,
 * particlebot_kernel_impl.cuh:541-831), bit for bit.
 *
 * Why this is allowed.  The force of robot j on robot i and the force of i on j are built from the
 * same operands with all signs flipped — rel = pos_j - pos_i, relvel = vel_j - vel_i, the unit vector
 * rel / dist — and every operation on the way (sub, fma, IEEE div and sqrt, the MUFU approximations of
 * gap^2) is sign-symmetric in round-to-nearest, while radA + radB, dist^2, dist, the gap, the normal
 * velocity rel.relvel and the |force| of a contact are invariant.  So F_ji = -F_ij EXACTLY, and a pair
 * needs one evaluation.  What is NOT free is the ORDER in which a robot adds its ~55 pair forces: fp32
 * addition does not associate, and the reference adds them in ascending slot order (rows -2..2 of the
 * stencil are ascending key ranges).  A robot's neighbours therefore split into a LOWER half (slots
 * before its own) and an UPPER half (slots after it):
 *
 *   phase A  every robot walks its lower half in order, adds the forces to its own sums directly (they
 *            come first in the reference's order) and PARKS each force in the partner's list at the
 *            position (rank) it has among that partner's upper neighbours;
 *   phase B  every robot subtracts its parked list in rank order (x - f == x + (-f) bit for bit), then
 *            the |force| of the parked contacts (a bit mask tells which entries are contacts; the norm is
 *            recomputed from the entry: squares do not see the sign).
 *
 * Both halves of a pair must live in the same block for this, so a block owns a 2-D PATCH of cells
 * (16 columns x up to 8 rows, ~280 robots at the density of the S1 workload).  The robots of the patch
 * and of the 2-cell halo around it are (rows + 4) contiguous slot ranges of the sorted array: one 1-D
 * TMA bulk copy each (cp.async.bulk + mbarrier complete_tx) brings their packed records into shared
 * memory, where they are spread into x / y / radius arrays kept twice (the second copy shifted by one
 * slot) so that ANY two consecutive neighbours arrive as one 64-bit load per array, already in the
 * halves of the packed fp32x2 registers the arithmetic works on.  Pairs with a halo robot are evaluated
 * from this side only (the owner of the halo robot does the same from its side): with 16 x 8 cells 79 %
 * of the ordered pairs are patch-internal, so 0.61 pair evaluations are left per ordered pair.
 *
 * A robot can receive at most PATCH_CAP parked forces; upper neighbours beyond that rank, and all halo
 * robots, are evaluated directly by the robot itself (nobody parks for them): exact for any density.
 * Patches that do not fit (more than PATCH_NT robots or PATCH_MAXREC staged records, a stencil that
 * wraps around the grid edge) and patches in which any pair left the admitted operand ranges of the
 * fast sequences (prs_collide.cuh) go through collide_robot, robot by robot.
 *
 * Only for fresh tables (the step just sorted: a robot's slot range IS its cell), plain swarms (no
 * transported object) and absForce_a not wanted; the launcher falls back to k_collide_exact otherwise.
 * Patches that hold robots are listed by the kernels that build the cell table (k_cell_apply,
 * k_reorder_packed): the grid is persistent (2 blocks per SM) and strides over that list.
 */
#pragma once
#include "prs_patchlist.cuh"

namespace prs {

constexpr int PATCH_SW = PATCH_W + 4;            /* staged cells per row (2-cell halo on both sides) */
constexpr int PATCH_HMAX = 8;                    /* patch rows, at most */
constexpr int PATCH_SHMAX = PATCH_HMAX + 4;
constexpr int PATCH_NT = 320;                    /* threads = robots a patch may hold */
constexpr int PATCH_MAXREC = 704;                /* staged records (patch + halo), at most */
constexpr int PATCH_CAP = 30;                    /* parked forces per robot, at most (<= 32: one mask word) */
constexpr int PATCH_LSTRIDE = 31;                /* list stride in float2 units (odd: bank spread) */
constexpr uint32_t PATCH_EMPTY = 0xffffffffu;

struct PatchSmem {
  /* parked forces, PATCH_LSTRIDE per robot; its first PATCH_MAXREC * 16 bytes are the TMA landing zone */
  float2 list[PATCH_NT * PATCH_LSTRIDE];
  float X[2][PATCH_MAXREC + 2], Y[2][PATCH_MAXREC + 2], R[2][PATCH_MAXREC + 2]; /* [1][m] = [0][m + 1] */
  uint32_t meta[PATCH_MAXREC + 2];                  /* per staged robot: rank gaps G1 | G2 << 16 (owned robots only) */
  uint32_t cmask[PATCH_NT];                      /* which parked entries are contacts */
  uint32_t pre[PATCH_SHMAX][PATCH_SW + 1];       /* per staged row: first record of cell c (dense), relative to the row */
  uint32_t rlo[PATCH_SHMAX], len[PATCH_SHMAX], base[PATCH_SHMAX + 1];   /* global slot of a row's first record, records, local index */
  uint32_t olo[PATCH_SHMAX], ocnt[PATCH_SHMAX], tbase[PATCH_SHMAX + 1]; /* owned part of a row: global slot, robots, thread index */
  uint32_t M, nown, fast, bad[2];
  unsigned long long bar;
};
static_assert(sizeof(float4) * PATCH_MAXREC <= sizeof(float2) * PATCH_NT * PATCH_LSTRIDE, "landing zone");
static_assert(PATCH_CAP <= 32 && PATCH_CAP < PATCH_LSTRIDE + 1 && (PATCH_MAXREC % 2) == 0, "list shape");
static_assert(2 * (sizeof(PatchSmem) + 1024) <= 227 * 1024, "two blocks per SM");

/* slow lane: one robot through the thread-per-robot code (global-memory neighbours, own cold path) */
__device__ __noinline__ void patch_fallback_robot(float2 *newVel, float *absForce_r, const float4 *pr, const float2 *svel,
                                                  const uint32_t *cellStart, const uint32_t *cellEnd, uint32_t k, float dt) {
  const PackedLayout in{pr, svel};
  collide_robot<false, false, PackedLayout>(newVel, nullptr, absForce_r, in, cellStart, cellEnd, k, dt);
}

enum { PATCH_PARK = 0, PATCH_STAGE = 1, PATCH_DIRECT = 2 };

__global__ void __launch_bounds__(PATCH_NT, 2)
k_collide_patch(float2 *__restrict__ newVel, float *absForce_r, const float4 *__restrict__ pr, const float2 *__restrict__ svel,
                const uint32_t *__restrict__ cellStart, const uint32_t *__restrict__ cellEnd,
                const uint32_t *__restrict__ patchList, const uint32_t *__restrict__ patchCount, uint32_t *nextCount, uint32_t PH,
                float dt, uint32_t *stats) {
  prs::pdl_sync();
  extern __shared__ __align__(128) unsigned char patch_smem_raw[];
  PatchSmem &S = *reinterpret_cast<PatchSmem *>(patch_smem_raw);
  const SimParams &P = c_prm.p;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t NW = PATCH_NT / 32;
  const int GX = (int)P.gridSize.x, GY = (int)P.gridSize.y;
  const uint32_t patches_x = (uint32_t)GX / PATCH_W;
  const uint32_t count = *patchCount;
  const uint32_t bar = smem_u32(&S.bar);
  if (tid == 0) { mbar_init(bar, 1); S.bad[0] = 0u; S.bad[1] = 0u; }
  if (blockIdx.x == 0 && tid == 0) *nextCount = 0u; /* the counter the NEXT sort step's table kernels append to */
  __syncthreads();

  /* per-robot constants of the pair arithmetic (plain swarm: one attraction product) */
  const float att_plain = __fmul_rn(1.0f, __fmul_rn(1.0f, P.attraction));
  const float spring_pos = P.spring;
  const float slope_plain = __fdiv_rn(__fadd_rn(__fdiv_rn(att_plain, __powf(0.0019f, 2.0f)), -2.5f), __fsub_rn(0.0019f, 0.0009f));
  const f32x2 DAMP2 = pk2(P.damping, P.damping), SHEAR2 = pk2(P.shear, P.shear), ATT2 = pk2(att_plain, att_plain);
  const f32x2 ZERO2 = pk2(0.0f, 0.0f), ONE2 = pk2(1.0f, 1.0f), HALF2 = pk2(0.5f, 0.5f);

  uint32_t it = 0, tma_phase = 0;
  for (uint32_t w = blockIdx.x; w < count; w += gridDim.x, it++) {
    const uint32_t p = patchList[w];
    const int cx0 = (int)(p % patches_x) * PATCH_W, cy0 = (int)(p / patches_x) * (int)PH;
    const int ph = min((int)PH, GY - cy0), sh = ph + 4;
    const bool edge = cx0 < 2 || cx0 + PATCH_W + 2 > GX || cy0 < 2 || cy0 + ph + 2 > GY;

    /* ---- tables: one warp per staged row reads its 20 cells; dense starts, row range, owned range ---- */
    for (int rr = (int)warp; rr < sh; rr += (int)NW) {
      uint32_t s = PATCH_EMPTY, e = 0u;
      if (lane < (uint32_t)PATCH_SW) {
        const uint32_t h = cell_hash(cx0 - 2 + (int)lane, cy0 - 2 + rr);
        s = __ldg(cellStart + h);
        const uint32_t e_raw = __ldg(cellEnd + h); /* stale for an empty cell: not used then */
        e = (s != PATCH_EMPTY) ? e_raw : 0u;
      }
      const unsigned ne = __ballot_sync(0xffffffffu, s != PATCH_EMPTY);
      uint32_t rlo = 0u, rhi = 0u, olo = 0u, ohi = 0u;
      if (ne) {
        rlo = __shfl_sync(0xffffffffu, s, __ffs(ne) - 1);
        rhi = __shfl_sync(0xffffffffu, e, 31 - __clz(ne));
      }
      const unsigned own_ne = ne & (((1u << PATCH_W) - 1u) << 2);
      if (own_ne) {
        olo = __shfl_sync(0xffffffffu, s, __ffs(own_ne) - 1);
        ohi = __shfl_sync(0xffffffffu, e, 31 - __clz(own_ne));
      }
      uint32_t ds = s; /* first record of the first non-empty cell at or after this one */
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, ds, o);
        if (lane + (uint32_t)o < 32u) ds = min(ds, t);
      }
      if (ds == PATCH_EMPTY) ds = rhi;
      if (lane <= (uint32_t)PATCH_SW) S.pre[rr][lane] = ds - rlo; /* lane PATCH_SW: the row's length */
      if (lane == 0) {
        S.rlo[rr] = rlo;
        S.len[rr] = rhi - rlo;
        S.olo[rr] = olo;
        S.ocnt[rr] = (rr >= 2 && rr < ph + 2) ? ohi - olo : 0u;
      }
    }
    __syncthreads();
    /* ---- local indices (prefix over the rows), can the patch be taken?, bulk copies ---- */
    if (warp == 0) {
      const uint32_t len = lane < (uint32_t)sh ? S.len[lane] : 0u, oc = lane < (uint32_t)sh ? S.ocnt[lane] : 0u;
      uint32_t il = len, io = oc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, io, o);
        if (lane >= (uint32_t)o) { il += a; io += b; }
      }
      const uint32_t M = __shfl_sync(0xffffffffu, il, 31), nown = __shfl_sync(0xffffffffu, io, 31);
      if (lane <= (uint32_t)sh) { S.base[lane] = il - len; S.tbase[lane] = io - oc; }
      const bool fast = !edge && M <= (uint32_t)PATCH_MAXREC && nown <= (uint32_t)PATCH_NT && nown > 0u;
      if (lane == 0) { S.M = M; S.nown = nown; S.fast = fast ? 1u : 0u; }
      if (fast) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the landing zone was written by plain stores before */
          mbar_expect_tx(bar, M * 16u);
        }
        __syncwarp();
        if (lane < (uint32_t)sh && len) tma_load_1d(smem_u32(S.list) + (il - len) * 16u, pr + S.rlo[lane], len * 16u, bar);
      }
    }
    __syncthreads();
    const uint32_t nown = S.nown;
    const bool has = tid < nown;
    /* which owned robot am I?  staged row r, global slot k, local index i_local */
    uint32_t r = 2u, k = 0u, i_local = 0u;
    if (has) {
      while (tid >= S.tbase[r + 1]) r++;
      const uint32_t q = tid - S.tbase[r];
      k = S.olo[r] + q;
      i_local = S.base[r] + (S.olo[r] - S.rlo[r]) + q;
    }
    if (stats && tid == 0) atomicAdd(&stats[S.fast ? 0 : 1], 1u); /* tuning aid: patches taken / handed to the slow lane */
    if (!S.fast) {
      for (uint32_t t = tid; t < nown; t += PATCH_NT) {
        uint32_t rr = 2u;
        while (t >= S.tbase[rr + 1]) rr++;
        patch_fallback_robot(newVel, absForce_r, pr, svel, cellStart, cellEnd, S.olo[rr] + (t - S.tbase[rr]), dt);
      }
      __syncthreads(); /* the tables are rewritten by the next patch */
      continue;
    }
    const float2 v_ = has ? __ldg(svel + k) : make_float2(0.0f, 0.0f);
    mbar_wait(bar, tma_phase & 1u);
    tma_phase++;

    /* ---- records -> x / y / radius arrays (twice, the second shifted by one slot); own record; rank gaps ---- */
    const uint32_t M = S.M;
    const float4 *land = reinterpret_cast<const float4 *>(S.list);
    for (uint32_t m = tid; m < M; m += PATCH_NT) {
      const float4 v = land[m];
      S.X[0][m] = v.x; S.Y[0][m] = v.y; S.R[0][m] = v.z;
      if (m) { S.X[1][m - 1] = v.x; S.Y[1][m - 1] = v.y; S.R[1][m - 1] = v.z; }
    }
    if (tid < 2u) { /* what a trip reads past the last record: anything finite (masked) */
      S.X[0][M + tid] = 0.0f; S.Y[0][M + tid] = 0.0f; S.R[0][M + tid] = 0.0f;
      S.X[1][M - 1u + tid] = 0.0f; S.Y[1][M - 1u + tid] = 0.0f; S.R[1][M - 1u + tid] = 0.0f;
      if (tid == 0u) { S.X[1][M + 1u] = 0.0f; S.Y[1][M + 1u] = 0.0f; S.R[1][M + 1u] = 0.0f; }
    }
    float px = 1.0f, py = 1.0f, rad = 0.0f;
    uint32_t orig = 0u;
    uint32_t lo[5] = {0u, 0u, 0u, 0u, 0u}, hi[5] = {0u, 0u, 0u, 0u, 0u};
    bool geometry_ok = true;
    if (has) {
      const float4 me = land[i_local];
      px = me.x; py = me.y; rad = me.z; orig = __float_as_uint(me.w);
      const int cxl = (((int)floorf((px - P.worldOrigin.x) / P.cellSize.x)) & (GX - 1)) - (cx0 - 2);
      geometry_ok = cxl >= 2 && cxl < PATCH_W + 2; /* a fresh table puts the robot in one of the patch's columns */
      if (geometry_ok) {
#pragma unroll
        for (int d = 0; d < 5; d++) {
          const uint32_t rr = r - 2u + (uint32_t)d;
          lo[d] = S.base[rr] + S.pre[rr][cxl - 2];
          hi[d] = S.base[rr] + S.pre[rr][cxl + 3];
        }
        geometry_ok = i_local >= lo[2] && i_local < hi[2];
        const uint32_t G1 = lo[3] - hi[2], G2 = G1 + lo[4] - hi[3];
        S.meta[i_local] = G1 | (G2 << 16);
      }
      S.cmask[tid] = 0u;
    }
    RangeAcc acc;
    acc.other = (att_admitted(att_plain) && geometry_ok) ? 0u : 1u;
    acc.robot(px, py);
    if (!geometry_ok) { lo[0] = lo[1] = lo[2] = lo[3] = lo[4] = 0u; hi[0] = hi[1] = hi[2] = hi[3] = hi[4] = 0u; i_local = 0u; }
    __syncthreads();

    float fx = 0.0f, fy = 0.0f;
    const float fr0 = has ? 0.0f * absForce_r[orig] : 0.0f; /* a NaN left there sticks, as in the reference (:688) */
    float fr = fr0;
    const f32x2 PX2 = pk2(px, px), PY2 = pk2(py, py), RAD2 = pk2(rad, rad), V2 = pk2(v_.x, v_.y);
    float2 *const mylist = S.list + tid * PATCH_LSTRIDE;

    /* One walk over the local indices [a, b), two neighbours per trip (an odd range ends with a masked
     * dummy).  PARK: lower neighbours — add to my sums, park the force for the partner if it is a patch
     * robot and the rank fits.  STAGE: upper neighbours nobody parks for (halo robots) — park the NEGATED
     * force in my own list at its rank.  DIRECT: upper neighbours beyond the list — add to my sums.
     *   row      staged row of the neighbours (ownership, global slots)
     *   g_shift / g_mask   which rank gap of the partner applies (PARK)
     *   rank_c   rank = rank_c + j for STAGE */
    auto walk = [&](auto mode_tag, uint32_t a, uint32_t b, uint32_t row, uint32_t g_shift, uint32_t g_mask, uint32_t rank_c) {
      constexpr int MODE = decltype(mode_tag)::value;
      if (b <= a) return;
      const uint32_t par = a & 1u;
      const float *XB = S.X[par] - par, *YB = S.Y[par] - par, *RB = S.R[par] - par;
      const uint32_t own_l = S.base[row] + (S.olo[row] - S.rlo[row]), own_n = S.ocnt[row];
      const uint32_t t_delta = S.tbase[row] - own_l;            /* partner's thread index = j + t_delta */
      const uint32_t slot_delta = S.rlo[row] - S.base[row];     /* partner's global slot = j + slot_delta */
      const uint32_t cm = i_local - 1u;
#pragma unroll 1
      for (uint32_t j = a; j < b; j += 2u) {
        const bool last = j + 1u == b;
        const f32x2 QX = *reinterpret_cast<const f32x2 *>(XB + j), QY = *reinterpret_cast<const f32x2 *>(YB + j);
        const f32x2 QR = *reinterpret_cast<const f32x2 *>(RB + j);
        const f32x2 RX = sub2(QX, PX2), RY = sub2(QY, PY2);
        const f32x2 D2 = fma2(RX, RX, mul2(RY, RY));
        float d20, d21;
        upk2(D2, d20, d21);
        acc.pair2(d20, last ? d20 : d21);
        /* the arithmetic of far2 / finish2 in prs_collide.cuh, operation for operation */
        const f32x2 Yv = pk2(rsqrt_approx(d20), rsqrt_approx(d21));
        const f32x2 Sq = mul2(D2, Yv), Hh = mul2(Yv, HALF2);
        const f32x2 DIST = fma2(fma2(sub2(ZERO2, Sq), Sq, D2), Hh, Sq);
        const f32x2 TOUCH = add2(RAD2, QR);
        const f32x2 ND = sub2(ZERO2, DIST);
        const f32x2 R1 = fma2(Yv, fma2(Yv, ND, ONE2), Yv);
        const f32x2 QXq = fma2(RX, R1, ZERO2), QYq = fma2(RY, R1, ZERO2);
        const f32x2 UX = fma2(R1, fma2(QXq, ND, RX), QXq), UY = fma2(R1, fma2(QYq, ND, RY), QYq);
        const f32x2 GAP = sub2(DIST, TOUCH);
        float g0, g1;
        upk2(GAP, g0, g1);
        const f32x2 Lg = pk2(lg2_approx(g0), lg2_approx(g1));
        const f32x2 L2 = add2(Lg, Lg);
        float l0, l1;
        upk2(L2, l0, l1);
        const float gg0 = ex2_approx(l0), gg1 = ex2_approx(l1);
        const f32x2 GG = pk2(gg0, gg1);
        const f32x2 NX = mul2(ATT2, UX), NY = mul2(ATT2, UY);
        const f32x2 RR0 = pk2(rcp_approx(gg0), rcp_approx(gg1));
        const f32x2 NG = sub2(ZERO2, GG);
        const f32x2 RR = fma2(RR0, fma2(RR0, NG, ONE2), RR0);
        const f32x2 TQX = fma2(NX, RR, ZERO2), TQY = fma2(NY, RR, ZERO2);
        const f32x2 TX = fma2(RR, fma2(TQX, NG, NX), TQX), TY = fma2(RR, fma2(TQY, NG, NY), TQY);
        float tx0, tx1, ty0, ty1;
        upk2(TX, tx0, tx1); upk2(TY, ty0, ty1);
        bool c0 = false, c1 = false;
        const bool n0 = g0 < 0.0019f, n1 = !last && g1 < 0.0019f;
        if (n0 || n1) { /* contacts and the two near-attraction regimes: rare per pair */
          float ux0, ux1, uy0, uy1;
          upk2(UX, ux0, ux1); upk2(UY, uy0, uy1);
          bool second = !n0;
#pragma unroll 1
          for (;;) {
            const float ux = second ? ux1 : ux0, uy = second ? uy1 : uy0, gap = second ? g1 : g0;
            float tx, ty;
            if (gap < 0.0f) { /* contact: dist < radA + radB */
              const f32x2 VB = __ldg(reinterpret_cast<const unsigned long long *>(svel + (j + slot_delta + (second ? 1u : 0u))));
              const f32x2 U = pk2(ux, uy);
              const f32x2 RV = sub2(VB, V2);
              float rvx, rvy;
              upk2(RV, rvx, rvy);
              const float dn = fmaf(uy, rvy, __fmul_rn(ux, rvx));
              const float ndn = -dn;
              const f32x2 TV = fma2(pk2(ndn, ndn), U, RV);
              const float sc = __fmul_rn(gap, spring_pos);
              f32x2 T = fma2(U, pk2(sc, sc), ZERO2);
              T = fma2(RV, DAMP2, T);
              T = fma2(SHEAR2, TV, T);
              upk2(T, tx, ty);
              if (MODE != PATCH_STAGE) {
                const float n2 = fmaf(tx, tx, __fmul_rn(ty, ty));
                acc.contact(n2);
                fr = __fadd_rn(fr, sqrt_fast_path(n2));
              }
              if (second) c1 = true; else c0 = true;
            } else {
              const float m = (gap < 0.0009f) ? 2.5f : fmaf(__fadd_rn(gap, -0.0009f), slope_plain, 2.5f);
              tx = __fmul_rn(ux, m);
              ty = __fmul_rn(uy, m);
            }
            if (second) { tx1 = tx; ty1 = ty; } else { tx0 = tx; ty0 = ty; }
            if (second || !n1) break;
            second = true;
          }
        }
        if (last) { tx1 = 0.0f; ty1 = 0.0f; }
        if (MODE == PATCH_STAGE) {
          const uint32_t rk = rank_c + j;
          mylist[rk] = make_float2(-tx0, -ty0);
          if (c0) S.cmask[tid] |= 1u << rk;
          if (!last) {
            mylist[rk + 1u] = make_float2(-tx1, -ty1);
            if (c1) S.cmask[tid] |= 1u << (rk + 1u);
          }
        } else {
          fx = __fadd_rn(__fadd_rn(fx, tx0), tx1);
          fy = __fadd_rn(__fadd_rn(fy, ty0), ty1);
          if (MODE == PATCH_PARK) {
            const uint32_t m0 = S.meta[j], m1 = S.meta[j + 1u];
            const uint32_t rk0 = cm - j - ((m0 >> g_shift) & g_mask), rk1 = cm - (j + 1u) - ((m1 >> g_shift) & g_mask);
            if (j - own_l < own_n && rk0 < (uint32_t)PATCH_CAP) {
              const uint32_t tj = j + t_delta;
              S.list[tj * PATCH_LSTRIDE + rk0] = make_float2(tx0, ty0);
              if (c0) atomicOr(&S.cmask[tj], 1u << rk0);
            }
            if (!last && j + 1u - own_l < own_n && rk1 < (uint32_t)PATCH_CAP) {
              const uint32_t tj = j + 1u + t_delta;
              S.list[tj * PATCH_LSTRIDE + rk1] = make_float2(tx1, ty1);
              if (c1) atomicOr(&S.cmask[tj], 1u << rk1);
            }
          }
        }
      }
    };
    using ParkT = std::integral_constant<int, PATCH_PARK>;
    using StageT = std::integral_constant<int, PATCH_STAGE>;
    using DirectT = std::integral_constant<int, PATCH_DIRECT>;

    /* ---- phase A: lower neighbours (rows -2, -1, my row before me), in the reference's order ---- */
    if (has) {
#pragma unroll 1
      for (int d = 0; d < 3; d++) {
        uint32_t a = lo[0], b = hi[0], gs = 16u, gm = 0xffffu;
        if (d == 1) { a = lo[1]; b = hi[1]; gs = 0u; }
        if (d == 2) { a = lo[2]; b = i_local; gm = 0u; }
        walk(ParkT{}, a, b, r - 2u + (uint32_t)d, gs, gm, 0u);
      }
    }
    __syncthreads();

    /* ---- phase B: upper neighbours.  Parked entries (rank < CAP, partner in the patch) are there;
     * halo neighbours of those ranks are evaluated now and parked by me; then everything is taken in
     * rank order; neighbours beyond the list are evaluated directly, in order, at the end. ---- */
    if (has) {
      const uint32_t up_a[3] = {i_local + 1u, lo[3], lo[4]}, up_b[3] = {hi[2], hi[3], hi[4]};
      const uint32_t G1 = lo[3] - hi[2], G2 = G1 + lo[4] - hi[3];
      const uint32_t gap_of[3] = {0u, G1, G2};
      uint32_t cut[3]; /* first local index of the segment whose rank does not fit the list */
      uint32_t total = 0u;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const uint32_t a = up_a[d], b = max(up_b[d], a);
        const uint32_t first_over = (uint32_t)PATCH_CAP + i_local + 1u + gap_of[d]; /* rank(j) = j - i - 1 - gap */
        cut[d] = max(a, min(b, first_over));
        total += b - a;
      }
#pragma unroll 1
      for (int d = 0; d < 3; d++) {
        uint32_t a = up_a[0], b = cut[0], g = 0u;
        if (d == 1) { a = up_a[1]; b = cut[1]; g = G1; }
        if (d == 2) { a = up_a[2]; b = cut[2]; g = G2; }
        if (b <= a) continue;
        const uint32_t row = r + (uint32_t)d;
        const uint32_t own_l = S.base[row] + (S.olo[row] - S.rlo[row]), own_h = own_l + S.ocnt[row];
        const uint32_t rank_c = 0u - i_local - 1u - g;
        const uint32_t left_b = min(b, max(a, own_l)), right_a = max(a, min(b, own_h));
        if (S.ocnt[row] == 0u) {
          walk(StageT{}, a, b, row, 0u, 0u, rank_c); /* a halo row: nobody parks */
        } else {
          walk(StageT{}, a, left_b, row, 0u, 0u, rank_c);
          walk(StageT{}, right_a, b, row, 0u, 0u, rank_c);
        }
      }
      const uint32_t parked = min(total, (uint32_t)PATCH_CAP);
      f32x2 F = pk2(fx, fy);
#pragma unroll 2
      for (uint32_t rk = 0; rk < parked; rk++) F = sub2(F, *reinterpret_cast<const f32x2 *>(mylist + rk));
      upk2(F, fx, fy);
      uint32_t cm_bits = S.cmask[tid];
      while (cm_bits) {
        const uint32_t rk = (uint32_t)__ffs(cm_bits) - 1u;
        cm_bits &= cm_bits - 1u;
        const float2 e = mylist[rk];
        const float n2 = fmaf(e.x, e.x, __fmul_rn(e.y, e.y));
        acc.contact(n2);
        fr = __fadd_rn(fr, sqrt_fast_path(n2));
      }
#pragma unroll 1
      for (int d = 0; d < 3; d++) {
        uint32_t a = cut[0], b = up_b[0];
        if (d == 1) { a = cut[1]; b = up_b[1]; }
        if (d == 2) { a = cut[2]; b = up_b[2]; }
        walk(DirectT{}, a, b, r + (uint32_t)d, 0u, 0u, 0u);
      }
      if (acc.outside(fx, fy, fr)) S.bad[it & 1u] = 1u;
    }
    __syncthreads();
    const bool bad = S.bad[it & 1u] != 0u;
    if (tid == 0) S.bad[(it + 1u) & 1u] = 0u;
    if (stats && tid == 0 && bad) atomicAdd(&stats[2], 1u);
    if (has) {
      if (bad) { /* some pair of the patch left the admitted ranges: both of its robots hold garbage */
        patch_fallback_robot(newVel, absForce_r, pr, svel, cellStart, cellEnd, k, dt);
      } else {
        v2 force = mk(fx, fy);
        const v2 pos = mk(px, py), vel = mk(v_.x, v_.y);
        obstacle_forces(pos, vel, rad, force, fr);
        const v2 nv = friction_and_velocity(vel, force, false, dt);
        newVel[orig] = make_float2(nv.x, nv.y);
        absForce_r[orig] = fr;
      }
    }
  }
}

}  // namespace prs
