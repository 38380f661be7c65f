// normal_ws.cu -- the one-pass D^dagger D kernel (two sites per thread, cp.async ring: normal.cu) with chunked
// self-scheduling instead of a static one-item-per-warp partition.
//
// Why: with the static partition ncu shows the SMs finishing between 339 K and 540 K active cycles (they do not
// get equal shares of the DRAM bandwidth) and the kernel ends when the slowest one does: SM-active is 85 % of the
// elapsed time, and unlike a streaming kernel the SMs that are still running cannot use the bandwidth the
// finished ones leave behind (the kernel is bound by per-warp latencies).
//
// Scheme: every strip (60 output columns) is cut into chunks of WS_CHUNK rows.  The static partition survives as
// the INITIAL ownership: warp w starts on segment w = (strip, row block), a run of consecutive chunks with a
// claim counter in global memory.  A warp claims its next chunk with one atomicAdd; when its own segment is
// used up it scans the counters, picks the segment with the most unclaimed chunks and claims from that one (the
// victim keeps claiming from the same counter, so nothing is lost or done twice).  Consecutive chunks of one
// strip continue the register windows and the cp.async ring without a prologue; a jump (stolen chunk, or a
// victim whose next chunk was taken) costs one prologue (4 input rows, 2 rows of D).
// Reproducibility: the reductions are accumulated PER CHUNK into a fixed slot of a global array and summed in slot
// order by the block that finishes last, so the result does not depend on which warp processed which chunk.
#include <cstdlib>
#include <type_traits>

#include "normal_args.cuh"

namespace glb {

constexpr int WS_THREADS = 128;
constexpr int WS_WARPS = WS_THREADS / 32;
constexpr int WS_OUT = 60;     // 64 loaded - 2 halo sites on each side
constexpr int WS_CHUNK = 30;   // rows per claim; a multiple of 3 keeps the register-ring phase at chunk starts

struct WsArgs {
  int* cnt;               // [nseg] chunks claimed so far from segment v (may overshoot its length)
  double* part;           // [nstrips * nchunks][3] per-chunk reduction partials
  unsigned int* ticket;   // blocks finished
  int nstrips, nchunks, nrb, nseg;
};

template <bool FUSE_XPAY, int NDOT, int STAGES>
__global__ void __launch_bounds__(WS_THREADS, 3) normal_ws_kernel(const NormArgs a, const WsArgs ws) {
  extern __shared__ __align__(32) unsigned char ring_raw[];
  double beta = 0.0;
  if (a.cg != nullptr) {
    if (a.cg->done) return;
    if (FUSE_XPAY) beta = xdiv(a.cg->rsq_new, a.cg->rsq_old);  // generic_cg.cpp:344
  }
  halo_wait_block(a.wait);
  constexpr int NRED = (NDOT == 0) ? 1 : (NDOT == 1 ? 2 : 3);
  constexpr int NARR = FUSE_XPAY ? 4 : 3;
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int X = a.X, Y = a.Y;
  const int nstrips = ws.nstrips, nchunks = ws.nchunks, nrb = ws.nrb, nseg = ws.nseg;
  const int wid = blockIdx.x * WS_WARPS + (threadIdx.x >> 5);
  const bool slab = (a.g_lo != nullptr);
  auto wrap_row = [&](int y) -> size_t { return (size_t)(((y % Y) + Y) % Y) * X; };
  auto link_row = [&](int y) -> ptrdiff_t { return (ptrdiff_t)y * X; };
  auto seg_begin = [&](int v) -> int { return (int)((long long)nchunks * (v / nstrips) / nrb); };
  auto seg_end = [&](int v) -> int { return (int)((long long)nchunks * (v / nstrips + 1) / nrb); };

  // ring slots (line-contiguous, odd rows swap neighbouring chunks: see normal.cu)
  auto phys = [](int j) -> int { return (j & ~7) | ((j & 7) ^ ((j >> 3) & 1)); };
  cplx* const ring_w = reinterpret_cast<cplx*>(ring_raw) + (size_t)(threadIdx.x >> 5) * 64;
  auto slot = [&](int stage, int arr) -> cplx* { return ring_w + (size_t)(stage * NARR + arr) * (WS_THREADS * 2); };
  const int ia0 = phys(lane), ia1 = phys(lane + 32);
  const int ir0 = phys(2 * lane), ir1 = phys(2 * lane + 1);

  // state that survives from one chunk to the next when they are consecutive rows of the same strip
  cplx p[3][2], t[3][2], ux[3][2], uy[3][2];
  cplx uxl[3];
  int strip = -1, next_row = -1;  // strip of the live windows and the row they are ready to produce
  int x0 = 0, xa = 0, xb = 0;
  bool active = false;
  int is_y = 0, is_st = 0, is_row2 = 0, is_lim = 0, rd_st = 0;

  auto issue = [&]() {  // copies for output row is_y: psi(is_y+2), U(is_y+1); always exactly one commit group
    if (is_y < is_lim) {
      const ptrdiff_t o1 = link_row(is_y + 1);
      if (slab && is_row2 >= Y) {
        const cplx* g = a.g_hi + (size_t)(is_row2 - Y) * X;
        cp_async16(slot(is_st, 0) + ia0, g + xa);
        cp_async16(slot(is_st, 0) + ia1, g + xb);
        if (FUSE_XPAY) {
          slot(is_st, 3)[ia0] = mk(0.0, 0.0);
          slot(is_st, 3)[ia1] = mk(0.0, 0.0);
        }
      } else {
        const size_t o2 = (size_t)is_row2 * X;
        const cplx* pa = (FUSE_XPAY ? a.r : a.in) + o2;
        cp_async16(slot(is_st, 0) + ia0, pa + xa);
        cp_async16(slot(is_st, 0) + ia1, pa + xb);
        if (FUSE_XPAY) {
          cp_async16(slot(is_st, 3) + ia0, a.pold + o2 + xa);
          cp_async16(slot(is_st, 3) + ia1, a.pold + o2 + xb);
        }
      }
      cp_async16(slot(is_st, 1) + ia0, a.Ux + o1 + xa);
      cp_async16(slot(is_st, 1) + ia1, a.Ux + o1 + xb);
      cp_async16(slot(is_st, 2) + ia0, a.Uy + o1 + xa);
      cp_async16(slot(is_st, 2) + ia1, a.Uy + o1 + xb);
      is_y++;
      if (++is_row2 == Y && !slab) is_row2 = 0;
      if (++is_st == STAGES) is_st = 0;
    }
    cp_async_commit();
  };
  auto ring_restart = [&](int y) {  // (re)start the prefetch pipeline at output row y
    cp_async_wait<0>();
    __syncwarp();
    is_y = y;
    is_st = 0;
    rd_st = 0;
    is_row2 = y + 2;
    if (!slab && is_row2 >= Y) is_row2 -= Y;
#pragma unroll
    for (int k = 0; k < STAGES - 1; k++) issue();
  };
  auto load_psi = [&](int y, cplx(&v)[2]) {
    if (slab && (y < 0 || y >= Y)) {
      ldv<2>((y < 0 ? a.g_lo + (size_t)(y + 2) * X : a.g_hi + (size_t)(y - Y) * X) + x0, v);
      return;
    }
    const size_t o = wrap_row(y) + x0;
    if (FUSE_XPAY) {
      cplx rr[2], pp[2];
      ldv<2>(a.r + o, rr);
      ldv<2>(a.pold + o, pp);
      v[0] = fadd(rr[0], fscale(beta, pp[0]));
      v[1] = fadd(rr[1], fscale(beta, pp[1]));
    } else {
      ldv<2>(a.in + o, v);
    }
  };
  auto prologue = [&](int s, int ya) {  // windows for output row ya of strip s: t(ya-1), t(ya) from psi(ya-2 .. ya+1)
    strip = s;
    const int xs = s * WS_OUT - 2 + 2 * lane;
    x0 = ((xs % X) + X) % X;
    active = (lane >= 1) && (lane <= 30) && (xs < X);
    const int win0 = s * WS_OUT - 2;
    xa = (((win0 + lane) % X) + X) % X;
    xb = (((win0 + 32 + lane) % X) + X) % X;
    cplx p_mm[2], p_m[2], ux_m[2], uy_mm[2];
    load_psi(ya - 2, p_mm);
    load_psi(ya - 1, p_m);
    load_psi(ya, p[0]);
    load_psi(ya + 1, p[1]);
    ldv_nc<2>(a.Uy + link_row(ya - 2) + x0, uy_mm);
    ldv_nc<2>(a.Ux + link_row(ya - 1) + x0, ux_m);
    ldv_nc<2>(a.Uy + link_row(ya - 1) + x0, uy[0]);
    ldv_nc<2>(a.Ux + link_row(ya) + x0, ux[1]);
    ldv_nc<2>(a.Uy + link_row(ya) + x0, uy[1]);
    const cplx uxl_m = shfl_up_c(ux_m[1], 1);
    uxl[1] = shfl_up_c(ux[1][1], 1);
    stag_row<false>(t[0], p_mm, p_m, p[0], ux_m, uxl_m, uy[0], uy_mm, a.mass);    // t(ya-1)
    stag_row<false>(t[1], p_m, p[0], p[1], ux[1], uxl[1], uy[1], uy[0], a.mass);  // t(ya)
  };
  auto row_step = [&](auto Kc, const int y) {
    constexpr int K0 = decltype(Kc)::value % 3, K1 = (K0 + 1) % 3, K2 = (K0 + 2) % 3;
    cplx la[2], lb[2];
    __syncwarp();
    issue();
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    la[0] = slot(rd_st, 0)[ir0];
    la[1] = slot(rd_st, 0)[ir1];
    ux[K2][0] = slot(rd_st, 1)[ir0];
    ux[K2][1] = slot(rd_st, 1)[ir1];
    uy[K2][0] = slot(rd_st, 2)[ir0];
    uy[K2][1] = slot(rd_st, 2)[ir1];
    if (FUSE_XPAY) {
      lb[0] = slot(rd_st, 3)[ir0];
      lb[1] = slot(rd_st, 3)[ir1];
      p[K2][0] = fadd(la[0], fscale(beta, lb[0]));
      p[K2][1] = fadd(la[1], fscale(beta, lb[1]));
    } else {
      p[K2][0] = la[0];
      p[K2][1] = la[1];
    }
    if (++rd_st == STAGES) rd_st = 0;
    uxl[K2] = shfl_up_c(ux[K2][1], 1);
    cplx res[2];
    stag_row<false>(t[K2], p[K0], p[K1], p[K2], ux[K2], uxl[K2], uy[K2], uy[K1], a.mass);  // t(y+1) = D psi
    stag_row<true>(res, t[K0], t[K1], t[K2], ux[K1], uxl[K1], uy[K1], uy[K0], a.mass);     // out(y) = D^dag t
    if (active) {
      const size_t o = (size_t)y * X + x0;
      stv<2>(a.out + o, res);
      if (FUSE_XPAY) stv<2>(a.pnew + o, p[K0]);
      if (NDOT >= 1) {
        cplx wv[2];
        if (a.w == nullptr) {
          wv[0] = p[K0][0];
          wv[1] = p[K0][1];
        } else {
          ldv<2>(a.w + o, wv);
        }
        Field<cplx>::dot_acc(acc, wv[0], res[0]);
        Field<cplx>::dot_acc(acc, wv[1], res[1]);
      }
      if (NDOT >= 2) {
        acc[2] += fnorm(res[0]);
        acc[2] += fnorm(res[1]);
      }
    }
  };

  // ---- claim loop
  int victim = (wid < nseg) ? wid : -1;
#pragma unroll 1
  for (;;) {
    int chunk = -1;
    while (chunk < 0) {
      if (victim >= 0) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&ws.cnt[victim], 1);
        idx = __shfl_sync(full, idx, 0);
        const int cb = seg_begin(victim), ce = seg_end(victim);
        if (cb + idx < ce) {
          chunk = cb + idx;
          break;
        }
      }
      // own segment (or the last victim) is used up: take from the segment with the most unclaimed chunks
      int best = 0, best_v = -1;
      for (int v = lane; v < nseg; v += 32) {
        const int len = seg_end(v) - seg_begin(v);
        const int rem = len - min(__ldcg(&ws.cnt[v]), len);
        if (rem > best) {
          best = rem;
          best_v = v;
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const int ob = __shfl_xor_sync(full, best, m), ov = __shfl_xor_sync(full, best_v, m);
        if (ob > best || (ob == best && ov > best_v)) {
          best = ob;
          best_v = ov;
        }
      }
      if (best <= 0) break;
      victim = best_v;
    }
    if (chunk < 0) break;

    const int s = victim % nstrips;
    const int ya = chunk * WS_CHUNK;
    const int yb = min(Y, ya + WS_CHUNK);
    const int lim = min(Y, seg_end(victim) * WS_CHUNK);  // prefetch may run ahead to the end of this segment
    if (s != strip || ya != next_row) {
      cp_async_wait<0>();
      __syncwarp();
      prologue(s, ya);
      is_lim = lim;
      ring_restart(ya);
    } else {
      is_lim = lim;
      if (is_y != ya + STAGES - 1) ring_restart(ya);  // the pipeline ran into its old limit: refill it
    }
    int y = ya;
#pragma unroll 1
    for (; y + 3 <= yb; y += 3) {
      row_step(std::integral_constant<int, 0>(), y);
      row_step(std::integral_constant<int, 1>(), y + 1);
      row_step(std::integral_constant<int, 2>(), y + 2);
    }
    if (y < yb) row_step(std::integral_constant<int, 0>(), y);
    if (y + 1 < yb) row_step(std::integral_constant<int, 1>(), y + 1);
    next_row = ((yb - ya) % 3 == 0) ? yb : -1;  // a short (last) chunk leaves the register-ring phase shifted
    if (NDOT > 0) {  // this chunk's partial sums -> its slot
#pragma unroll
      for (int r = 0; r < NRED; r++) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc[r] += shfl_xor_d(acc[r], m);
        if (lane == 0) ws.part[((size_t)s * nchunks + chunk) * 3 + r] = acc[r];
        acc[r] = 0.0;
      }
    }
  }
  cp_async_wait<0>();

  // ---- the block that finishes last sums the chunk partials in slot order and re-arms the counters
  __shared__ bool s_last;
  __shared__ double s_red[NRED * 32];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicInc(ws.ticket, gridDim.x - 1);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int v = threadIdx.x; v < nseg; v += blockDim.x) ws.cnt[v] = 0;
  if (NDOT > 0) {
    double tot[NRED];
    const int nslots = nstrips * nchunks;
#pragma unroll
    for (int r = 0; r < NRED; r++) {
      tot[r] = 0.0;
      for (int i = threadIdx.x; i < nslots; i += blockDim.x) tot[r] += __ldcg(&ws.part[(size_t)i * 3 + r]);
    }
    block_sum<NRED>(tot, s_red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int r = 0; r < NRED; r++) {
        a.red.result_dev[r] = tot[r];
        if (a.red.result_host) a.red.result_host[r] = tot[r];
      }
    }
    // slab run over peer memory: the first warp finishes the sum over ranks
    if (a.cg != nullptr && a.cg_role == 3 && threadIdx.x < 32) p2p_allreduce_warp(a.pr, tot, 2);
    if (threadIdx.x == 0 && a.cg != nullptr) {
      if (a.cg_role == 1 || a.cg_role == 3) {
        a.cg->pAp_re = tot[0];
        a.cg->pAp_im = tot[1];
        a.cg->rsq_old = a.cg->rsq_new;
      } else if (a.cg_role == 2) {
        a.cg->partial[1] = tot[0];
        a.cg->partial[2] = tot[1];
      }
    }
  }
}

template <bool FUSE, int NDOT, int STAGES>
static int launch_ws_t(glb_operator* op, const NormArgs& a) {
  glb_context* ctx = op->ctx;
  auto kern = normal_ws_kernel<FUSE, NDOT, STAGES>;
  const size_t smem = (size_t)STAGES * (FUSE ? 4 : 3) * WS_THREADS * 2 * sizeof(cplx);
  static int per_sm = 0;
  if (per_sm == 0) {
    if (smem + 4096 > 48 * 1024) GLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WS_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
  }
  WsArgs ws;
  ws.nstrips = (a.X + WS_OUT - 1) / WS_OUT;
  ws.nchunks = (a.Y + WS_CHUNK - 1) / WS_CHUNK;
  const long long max_warps = (long long)ctx->sm_count * per_sm * WS_WARPS;
  long long nrb = max_warps / ws.nstrips;
  if (nrb > ws.nchunks) nrb = ws.nchunks;
  if (nrb < 1) nrb = 1;
  ws.nrb = (int)nrb;
  ws.nseg = ws.nstrips * ws.nrb;
  // workspace (grown on demand, zeroed once: the kernel re-arms it itself)
  const size_t need_cnt = (size_t)ws.nseg, need_part = (size_t)ws.nstrips * ws.nchunks * 3;
  if (need_cnt > ctx->ws_cnt_n || need_part > ctx->ws_part_n || !ctx->ws_ticket) {
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->ws_cnt);
    cudaFree(ctx->ws_part);
    if (!ctx->ws_ticket) {
      GLB_CUDA(cudaMalloc((void**)&ctx->ws_ticket, sizeof(unsigned int)));
      GLB_CUDA(cudaMemset(ctx->ws_ticket, 0, sizeof(unsigned int)));
    }
    ctx->ws_cnt_n = need_cnt * 2;
    ctx->ws_part_n = need_part * 2;
    GLB_CUDA(cudaMalloc((void**)&ctx->ws_cnt, sizeof(int) * ctx->ws_cnt_n));
    GLB_CUDA(cudaMemset(ctx->ws_cnt, 0, sizeof(int) * ctx->ws_cnt_n));
    GLB_CUDA(cudaMalloc((void**)&ctx->ws_part, sizeof(double) * ctx->ws_part_n));
  }
  ws.cnt = ctx->ws_cnt;
  ws.part = ctx->ws_part;
  ws.ticket = ctx->ws_ticket;
  long long blocks = (ws.nseg + WS_WARPS - 1) / WS_WARPS;
  // every resident slot gets a block even when there are fewer initial segments: the extra warps start by stealing
  if (blocks < (long long)ctx->sm_count * per_sm) blocks = (long long)ctx->sm_count * per_sm;
  if (blocks > (long long)ctx->sm_count * per_sm) blocks = (long long)ctx->sm_count * per_sm;
  ProfScope prof(ctx, FUSE ? PROF_NORMAL_FUSED : PROF_NORMAL);
  kern<<<(unsigned)blocks, WS_THREADS, smem, ctx->stream>>>(a, ws);
  GLB_LAUNCH_CHECK();
  return GLB_OK;
}

int launch_normal_ws(glb_operator* op, const NormArgs& a, bool fuse, int ndot) {
  if (fuse) {
    if (ndot == 0) return launch_ws_t<true, 0, 4>(op, a);
    if (ndot == 1) return launch_ws_t<true, 1, 4>(op, a);
    return launch_ws_t<true, 2, 4>(op, a);
  }
  if (ndot == 0) return launch_ws_t<false, 0, 4>(op, a);
  if (ndot == 1) return launch_ws_t<false, 1, 4>(op, a);
  return launch_ws_t<false, 2, 4>(op, a);
}

}  // namespace glb
