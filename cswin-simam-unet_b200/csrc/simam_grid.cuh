// SimAM, token layout (B, L, C): GRID-RESIDENT kernels (included by simam.cu inside its anonymous namespace).
//
// The cluster kernels above give an image to one cluster of <= 8 CTAs: at batch 32 only 128 of the 148 SMs
// hold a cluster, every CTA runs "read everything -> exchange -> write everything" in lock-step with all the
// others (reads and writes never overlap) and the second sweep re-reads the image through L2 (1.35x DRAM reads
// in the backward pass, whose x + grad_y do not fit).  Here the WHOLE GPU works on a few images at a time:
//
//   * the batch is cut into rounds of k whole images and every image into m pieces of token rows, k * m <= #SMs:
//     one piece per CTA per round (one persistent CTA per SM), so a CTA's share of a round —
//     <= ~70 KB — stays in SHARED MEMORY between the statistics and the rescale: every byte crosses HBM once;
//   * four share buffers, all traffic on the copy engine (cp.async.bulk both ways, the rescale runs in place in
//     shared memory): while round r - 1 is being written out, round r waits for its statistics, round r + 1 has
//     landed and round r + 2 is loading, so DRAM sees reads and writes at the same time and no thread ever
//     stalls on a store;
//   * per-(image, channel) moments: every CTA writes the partial sums of its rows to its own slot of a
//     workspace and bumps the image's counter; once the counter says every contributor has published, each CTA
//     sums the slots IN SLOT ORDER (bit-identical in every CTA, deterministic from run to run);
//   * the workspace cleans itself: the last CTA to read an image's slots zeroes its two counters, so the caller
//     zeroes the workspace once, when it allocates it.
//
// A share starts on a token-row boundary and 512 % (vectors per row) == 0, so a thread meets the same 8 (bf16) /
// 4 (fp32) channels in every vector it handles and keeps their moments in registers, exactly like the
// streaming kernels.  Every image is cut the same way (m pieces of ceil(L / m) rows, one piece per CTA), so the
// order of every sum depends on the image alone: results are bit-identical under batch permutation.

#ifdef CSB_PROF
__device__ unsigned long long g_simam_prof[16];
#define GR_T(v) const long long v = clock64()
#define GR_ADD(i, a, b) if (blockIdx.x == 1 && threadIdx.x == 0) g_simam_prof[i] += (unsigned long long)((b) - (a))
#else
#define GR_T(v)
#define GR_ADD(i, a, b)
#endif

constexpr int GR_THREADS = 512;
constexpr int GR_NBUF = 4;
constexpr int GR_NSL = 6;   // slot loads per thread when an image's partial moments are summed
constexpr int GR_MAX_SMEM = 227 * 1024;
constexpr int GR_MAX_BATCH = 4096;
constexpr int GR_COUNTER_BYTES = 2 * GR_MAX_BATCH * 4;

struct GridPlan {
  int B, L;        // images, token rows per image
  int nround;      // rounds; round r holds B / nround (+1 for r < B % nround) whole images
  int m, rpi;      // every image is cut into m pieces of rpi = ceil(L / m) token rows: one piece per CTA per round
  int buf_bytes;   // one shared-memory buffer (backward: the x share, then the grad_y share at buf_bytes / 2)
};

template <int VE, int CVEC>
struct GrSmem {
  static constexpr int GW = CVEC < 32 ? CVEC : 32;        // column vectors a warp covers
  static constexpr int NCLS = CVEC < 32 ? 1 : CVEC / 32;  // warp w covers vectors (w % NCLS) * 32 + lane
  static constexpr int CW = CVEC * VE;                    // channels
  static constexpr int WARPS = GR_THREADS / 32;
  static constexpr int RED_RAW = WARPS * 2 * VE * GW * 4;    // [WARPS][2][VE][GW] floats
  static constexpr int RED_BYTES = RED_RAW < 8192 ? 8192 : RED_RAW;  // also [GROUPS][2 CW] floats = 8 KB (totals)
  static constexpr int NQ = 2 * CW / 4;                      // float4 quads of one slot
  static constexpr int GROUPS = GR_THREADS / NQ;             // slot groups when the CTA reads an image's slots
  static constexpr int FIN_BYTES = 2 * CW * 4;               // [2][CW] floats, twice (own partial, image total)
  static constexpr int TAIL_BYTES = RED_BYTES + 2 * FIN_BYTES + 64;
  static constexpr int MAX_BUF = ((GR_MAX_SMEM - TAIL_BYTES) / GR_NBUF) / 256 * 256;
};

struct GrRound {
  int b0, k;  // first image, images
};
__device__ __forceinline__ GrRound gr_round(const GridPlan& p, int r) {
  GrRound g;
  const int base = p.B / p.nround, extra = p.B % p.nround;
  g.k = base + (r < extra ? 1 : 0);
  g.b0 = r * base + (r < extra ? r : extra);
  return g;
}
// This CTA's share of a round: rows [row0, row0 + nrows) of image img (nrows == 0: nothing this round).  The cut
// depends on the image alone, never on its place in the batch: results are invariant under batch permutation.
struct GrPart {
  int img, row0, nrows, slot, expected;
};
__device__ __forceinline__ GrPart gr_part(const GridPlan& p, const GrRound& g, int cta) {
  GrPart pt;
  const int j = cta / p.m, piece = cta % p.m;
  pt.img = g.b0 + j;
  pt.row0 = piece * p.rpi;
  const int left = p.L - pt.row0;
  pt.nrows = (j < g.k && left > 0) ? (left < p.rpi ? left : p.rpi) : 0;
  pt.slot = piece;
  pt.expected = (p.L + p.rpi - 1) / p.rpi;  // non-empty pieces of an image
  return pt;
}

// acc[q][e] summed over the threads of the CTA that hold the same channel -> s_out[q * CW + channel]; fixed order
template <int VE, int CVEC>
__device__ __forceinline__ void gr_block_reduce(float (&acc)[2][VE], float* s_red, float* s_out) {
  using S = GrSmem<VE, CVEC>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int e = 0; e < VE; ++e) {
#pragma unroll
      for (int o = 16; o >= CVEC; o >>= 1) acc[q][e] += __shfl_xor_sync(0xffffffffu, acc[q][e], o);
    }
  if (lane < S::GW) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < VE; ++e) s_red[((warp * 2 + q) * VE + e) * S::GW + lane] = acc[q][e];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * S::CW; j += GR_THREADS) {
    const int q = j / S::CW, ch = j % S::CW, g = ch / VE, e = ch % VE;
    const int cls = g / S::GW, gl = g % S::GW;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < S::WARPS / S::NCLS; ++w) a += s_red[(((cls + w * S::NCLS) * 2 + q) * VE + e) * S::GW + gl];
    s_out[j] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ int gr_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int CVEC, bool BWD>
__global__ void __launch_bounds__(GR_THREADS, 1)
    simam_nlc_grid(const T* __restrict__ x, const T* __restrict__ gy, const float* __restrict__ stats_in,
                   float* __restrict__ stats_out, T* __restrict__ out, const GridPlan p, float* __restrict__ gpart,
                   int* __restrict__ count, int* __restrict__ done, float e_lambda) {
  constexpr int VE = Vec16<T>::N, NP = VE / 2;
  constexpr bool BF = sizeof(T) == 2;
  using S = GrSmem<VE, CVEC>;
  constexpr int CW = S::CW;
  constexpr int ROW_BYTES = CW * (int)sizeof(T);
  extern __shared__ __align__(128) uint8_t gr_smem[];
  uint8_t* bufs = gr_smem;
  float* s_red = reinterpret_cast<float*>(gr_smem + GR_NBUF * p.buf_bytes);
  float* s_fin = s_red + S::RED_BYTES / 4;
  float* s_tot = s_fin + 2 * CW;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_tot + 2 * CW);
  const int cta = (int)blockIdx.x;
  const int tid = threadIdx.x, cv = tid % CVEC;
  const int half = p.buf_bytes / 2;  // backward: offset of the grad_y share inside a buffer
  const float Lf = (float)p.L;

  if (tid == 0) {
    for (int i = 0; i < GR_NBUF; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&full[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- loads: thread 0 moves this CTA's share of round r into buffer r % GR_NBUF ----
  auto issue_load = [&](int r) {
    if (r >= p.nround) return;
    const GrPart pt = gr_part(p, gr_round(p, r), cta);
    if (pt.nrows == 0) return;
    const uint32_t bytes = (uint32_t)pt.nrows * ROW_BYTES;
    const int64_t off = ((int64_t)pt.img * p.L + pt.row0) * ROW_BYTES;
    const uint32_t bar = st_smem_u32(&full[r % GR_NBUF]);
    const uint32_t dst = st_smem_u32(bufs + (r % GR_NBUF) * p.buf_bytes);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer's last readers were generic loads
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BWD ? 2 * bytes : bytes) : "memory");
    constexpr uint32_t PIECE = 32768;
    for (uint32_t o = 0; o < bytes; o += PIECE) {
      const uint32_t n = bytes - o < PIECE ? bytes - o : PIECE;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + o),
                   "l"(reinterpret_cast<const uint8_t*>(x) + off + o), "r"(n), "r"(bar)
                   : "memory");
      if constexpr (BWD)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + half + o),
                     "l"(reinterpret_cast<const uint8_t*>(gy) + off + o), "r"(n), "r"(bar)
                     : "memory");
    }
  };

  // ---- stores: thread 0 hands the rescaled share of round r to the copy engine (cp.async.bulk shared -> global).
  //      The threads never stall on store back-pressure, and the write-out of round r - 1 overlaps the loads of
  //      round r + 2 in the memory system (thread-issued stores kept every SM in a pure-write phase for ~2 us per
  //      round while no load was in flight). ----
  auto issue_store = [&](int r) {
    const GrPart pt = gr_part(p, gr_round(p, r), cta);
    if (pt.nrows > 0) {
      const uint32_t bytes = (uint32_t)pt.nrows * ROW_BYTES;
      const int64_t off = ((int64_t)pt.img * p.L + pt.row0) * ROW_BYTES;
      const uint32_t src = st_smem_u32(bufs + (r % GR_NBUF) * p.buf_bytes);
      const uint64_t pol_first = l2_evict_first_policy();
      constexpr uint32_t PIECE = 32768;
      for (uint32_t o = 0; o < bytes; o += PIECE) {
        const uint32_t n = bytes - o < PIECE ? bytes - o : PIECE;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                         reinterpret_cast<uint8_t*>(out) + off + o),
                     "r"(src + o), "r"(n), "l"(pol_first)
                     : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // one group per round, empty or not
  };

  uint32_t use_par = 0;  // bit b: parity of the next completion of full[b] this CTA waits for

  // ---- first sweep: partial moments of this CTA's rows, in registers ----
  auto accumulate = [&](int r, const GrPart& pt, float (&acc)[2][VE]) {
    const uint8_t* buf = bufs + (r % GR_NBUF) * p.buf_bytes;
    const uint4* vx = reinterpret_cast<const uint4*>(buf);
    const uint4* vg = reinterpret_cast<const uint4*>(buf + half);
    float pivot[VE];  // pivot = row 0 of the image, this thread's channels
    const uint4* ximg = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(x) + (int64_t)pt.img * p.L * ROW_BYTES);
    unpack<T>(__ldg(ximg + cv), pivot);
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
    const int v_hi = pt.nrows * CVEC;
    int v = tid;
    if constexpr (!BWD) {
      if constexpr (BF) {
        f2_t npivot2[NP], sum2[NP], sq2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          npivot2[q] = f2_make(-pivot[2 * q], -pivot[2 * q + 1]);
          sum2[q] = sq2[q] = f2_splat(0.f);
        }
        for (; v < v_hi; v += GR_THREADS) {
          const uint4 u = vx[v];
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f2_t d = f2_add(f2_from_bf16x2(w[q]), npivot2[q]);
            sum2[q] = f2_add(sum2[q], d);
            sq2[q] = f2_fma(d, d, sq2[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          f2_split(sum2[q], acc[0][2 * q], acc[0][2 * q + 1]);
          f2_split(sq2[q], acc[1][2 * q], acc[1][2 * q + 1]);
        }
      } else {
        for (; v < v_hi; v += GR_THREADS) {
          float f[VE];
          unpack<T>(vx[v], f);
#pragma unroll
          for (int e = 0; e < VE; ++e) {
            const float d = f[e] - pivot[e];
            acc[0][e] += d;
            acc[1][e] = fmaf(d, d, acc[1][e]);
          }
        }
      }
    } else {
      const int64_t p0 = (int64_t)pt.img * CW + cv * VE;
      float dmean[VE], inv4v[VE];  // mean - pivot and 1 / (4 variance) from the forward pass
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        dmean[e] = __ldg(stats_in + 2 * (p0 + e));
        inv4v[e] = 1.f / (4.f * __ldg(stats_in + 2 * (p0 + e) + 1));
      }
      if constexpr (BF) {
        const f2_t quarter2 = f2_splat(0.25f), mone2 = f2_splat(-1.f);
        f2_t r1p[NP], r2p[NP], nmean2[NP], inv8v2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          r1p[q] = r2p[q] = f2_splat(0.f);
          nmean2[q] = f2_make(-(pivot[2 * q] + dmean[2 * q]), -(pivot[2 * q + 1] + dmean[2 * q + 1]));
          inv8v2[q] = f2_make(0.5f * inv4v[2 * q], 0.5f * inv4v[2 * q + 1]);
        }
        for (; v < v_hi; v += GR_THREADS) {
          const uint4 ux = vx[v], ug = vg[v];
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
            const f2_t t = f2_add(x2, nmean2[q]), dd = f2_mul(t, t);
            const f2_t th = f2_tanh(f2_fma(dd, inv8v2[q], quarter2));
            const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));  // -4a = g x (tanh^2 - 1)
            r1p[q] = f2_fma(na4, dd, r1p[q]);
            r2p[q] = f2_fma(na4, t, r2p[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          float lo, hi;
          f2_split(r1p[q], lo, hi);
          acc[0][2 * q] = -lo;
          acc[0][2 * q + 1] = -hi;
          f2_split(r2p[q], lo, hi);
          acc[1][2 * q] = -lo;
          acc[1][2 * q + 1] = -hi;
        }
      } else {
        for (; v < v_hi; v += GR_THREADS) {
          float fx[VE], fg[VE];
          unpack<T>(vx[v], fx);
          unpack<T>(vg[v], fg);
#pragma unroll
          for (int e = 0; e < VE; ++e) {
            const float t = centred(fx[e], pivot[e], dmean[e]), dd = t * t;
            const float sg_ = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
            const float a4 = 4.f * fg[e] * fx[e] * sg_ * (1.f - sg_);
            acc[0][e] = fmaf(a4, dd, acc[0][e]);
            acc[1][e] = fmaf(a4, t, acc[1][e]);
          }
        }
      }
    }
  };
  // ---- ... -> this CTA's slot of the image, then the image's counter + 1 (release: fire and forget) ----
  auto publish = [&](const GrPart& pt, float (&acc)[2][VE]) {
    gr_block_reduce<VE, CVEC>(acc, s_red, s_fin);
    float* slot = gpart + ((int64_t)pt.img * p.m + pt.slot) * (2 * CW);
    for (int j = tid; j < 2 * CW; j += GR_THREADS) __stcg(slot + j, s_fin[j]);
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(count + pt.img) : "memory");
  };
  // ---- the image's moments: wait for every contributor, read the slots (one L2 round trip: thread = (slot
  //      group, float4 quad), <= GR_NSL loads each, all in flight together) ... ----
  auto wait_published = [&](const GrPart& pt, int seen) {
    if ((tid & 31) == 0) {
      unsigned spins = 0;
      while (seen < pt.expected) {
        seen = gr_ld_acquire(count + pt.img);
        if (++spins > (1u << 28)) __trap();  // a dirty workspace (see the header) would otherwise hang the GPU
      }
    }
    __syncwarp();
  };
  auto fetch_slots = [&](const GrPart& pt, float4 (&sl)[GR_NSL]) {
    const int quad = tid % S::NQ, grp = tid / S::NQ;
    const float4* base = reinterpret_cast<const float4*>(gpart + (int64_t)pt.img * p.m * (2 * CW)) + quad;
#pragma unroll
    for (int i = 0; i < GR_NSL; ++i) {
      const int k = grp + i * S::GROUPS;
      sl[i] = k < pt.expected ? __ldcg(base + (int64_t)k * S::NQ) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // ---- ... and their sum in a fixed order (slots grp, grp + GROUPS, ... per thread, then the groups in order):
  //      bit-identical in every CTA and from run to run ----
  auto totals = [&](const GrPart& pt, const float4 (&sl)[GR_NSL]) {
    float4 a = sl[0];
#pragma unroll
    for (int i = 1; i < GR_NSL; ++i) {
      a.x += sl[i].x;
      a.y += sl[i].y;
      a.z += sl[i].z;
      a.w += sl[i].w;
    }
    reinterpret_cast<float4*>(s_red)[tid] = a;  // [grp][quad]
    __syncthreads();
    for (int j = tid; j < 2 * CW; j += GR_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < S::GROUPS; ++g) t += s_red[g * (2 * CW) + j];
      s_tot[j] = t;
    }
    __syncthreads();
    if (tid == 0) {  // the last reader leaves the image's counters zeroed for the next call
      if (atomicAdd(done + pt.img, 1) == pt.expected - 1) {
        count[pt.img] = 0;
        done[pt.img] = 0;
      }
    }
  };

  // ---- second sweep: rescale this CTA's rows with the image's moments (s_tot), in place in shared memory ----
  auto rescale = [&](int r, const GrPart& pt) {
    uint8_t* buf = bufs + (r % GR_NBUF) * p.buf_bytes;
    const uint4* vx = reinterpret_cast<const uint4*>(buf);
    const uint4* vg = reinterpret_cast<const uint4*>(buf + half);
    uint4* vout = reinterpret_cast<uint4*>(buf);  // IN PLACE: the copy engine writes the share out (issue_store)
    float pivot[VE];
    unpack<T>(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(x) + (int64_t)pt.img * p.L * ROW_BYTES) + cv),
              pivot);
    const int v_hi = pt.nrows * CVEC;
    int v = tid;
    float t0[VE], t1[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      t0[e] = s_tot[cv * VE + e];
      t1[e] = s_tot[CW + cv * VE + e];
    }
    if constexpr (!BWD) {
      float dmean[VE], var[VE];
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        dmean[e] = t0[e] / Lf;
        var[e] = fmaxf(t1[e] - t0[e] * dmean[e], 0.f) / (Lf - 1.f) + e_lambda;
      }
      if (stats_out != nullptr && pt.slot == 0 && tid < CVEC) {
        const int64_t p0 = (int64_t)pt.img * CW + cv * VE;
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          stats_out[2 * (p0 + e)] = dmean[e];
          stats_out[2 * (p0 + e) + 1] = var[e];
        }
      }
      if constexpr (BF) {
        const f2_t quarter2 = f2_splat(0.25f), half2 = f2_splat(0.5f);
        f2_t nmean2[NP], inv8v2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          nmean2[q] = f2_make(-(pivot[2 * q] + dmean[2 * q]), -(pivot[2 * q + 1] + dmean[2 * q + 1]));
          inv8v2[q] = f2_make(1.f / (8.f * var[2 * q]), 1.f / (8.f * var[2 * q + 1]));
        }
        for (; v < v_hi; v += GR_THREADS) {
          const uint4 u = vx[v];
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // y = hx tanh(t^2 / (8v) + 1/4) + hx, hx = x / 2
            const f2_t x2 = f2_from_bf16x2(w[q]);
            const f2_t t = f2_add(x2, nmean2[q]);
            const f2_t th = f2_tanh(f2_fma(f2_mul(t, t), inv8v2[q], quarter2));
            const f2_t hx = f2_mul(x2, half2);
            float lo, hi;
            f2_split(f2_fma(hx, th, hx), lo, hi);
            o[q] = pack_bf16x2(lo, hi);
          }
          vout[v] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
        float inv4v[VE];
#pragma unroll
        for (int e = 0; e < VE; ++e) inv4v[e] = 1.f / (4.f * var[e]);
        for (; v < v_hi; v += GR_THREADS) {
          float f[VE];
          unpack<T>(vx[v], f);
#pragma unroll
          for (int e = 0; e < VE; ++e) f[e] = simam_fwd_elem<T>(f[e], FwdCoef{pivot[e], dmean[e], inv4v[e]});
          vout[v] = pack<T>(f);
        }
      }
    } else {
      const int64_t p0 = (int64_t)pt.img * CW + cv * VE;
      float dmean[VE], vv[VE], inv4v[VE], c1[VE], c2[VE];
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        dmean[e] = __ldg(stats_in + 2 * (p0 + e));
        vv[e] = __ldg(stats_in + 2 * (p0 + e) + 1);
        inv4v[e] = 1.f / (4.f * vv[e]);
        const float r1 = 0.25f * t0[e], r2 = 0.25f * t1[e];  // the sweeps accumulate 4 a
        c1[e] = r1 * inv4v[e] / (vv[e] * (Lf - 1.f));         // R1 / (4 v^2 n)
        c2[e] = 2.f / Lf * r2 * inv4v[e];                     // (2/L) R2
      }
      if constexpr (BF) {
        const f2_t quarter2 = f2_splat(0.25f), mone2 = f2_splat(-1.f), half2 = f2_splat(0.5f);
        f2_t nmean2[NP], inv8v2[NP], nk2[NP], nc2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          nmean2[q] = f2_make(-(pivot[2 * q] + dmean[2 * q]), -(pivot[2 * q + 1] + dmean[2 * q + 1]));
          inv8v2[q] = f2_make(0.5f * inv4v[2 * q], 0.5f * inv4v[2 * q + 1]);
          nk2[q] = f2_make(-2.f * c1[2 * q], -2.f * c1[2 * q + 1]);
          nc2[q] = f2_make(-c2[2 * q], -c2[2 * q + 1]);
        }
        for (; v < v_hi; v += GR_THREADS) {
          const uint4 ux = vx[v], ug = vg[v];
          const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
          uint32_t o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // grad_x = 0.5 g (1 + tanh) + t (4a k1 - k2) - c2 with k1 = 1/(8v)
            const f2_t x2 = f2_from_bf16x2(wx[q]), g2 = f2_from_bf16x2(wg[q]);
            const f2_t t = f2_add(x2, nmean2[q]), dd = f2_mul(t, t);
            const f2_t th = f2_tanh(f2_fma(dd, inv8v2[q], quarter2));
            const f2_t na4 = f2_mul(f2_mul(g2, x2), f2_fma(th, th, mone2));
            const f2_t hg = f2_mul(g2, half2);
            const f2_t gs = f2_fma(hg, th, f2_add(hg, nc2[q]));
            const f2_t res = f2_fma(t, f2_fma(f2_mul(na4, mone2), inv8v2[q], nk2[q]), gs);
            float lo, hi;
            f2_split(res, lo, hi);
            o[q] = pack_bf16x2(lo, hi);
          }
          vout[v] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      } else {
        for (; v < v_hi; v += GR_THREADS) {
          float fx[VE], fg[VE];
          unpack<T>(vx[v], fx);
          unpack<T>(vg[v], fg);
#pragma unroll
          for (int e = 0; e < VE; ++e) {
            const float t = centred(fx[e], pivot[e], dmean[e]), dd = t * t;
            const float sg_ = Sig<T>::f(fmaf(dd, inv4v[e], 0.5f));
            const float a = fg[e] * fx[e] * sg_ * (1.f - sg_);
            fx[e] = fmaf(fg[e], sg_, 2.f * t * fmaf(a, inv4v[e], -c1[e])) - c2[e];
          }
          vout[v] = pack<T>(fx);
        }
      }
    }
  };

  if (tid == 0)
    for (int r = 0; r < GR_NBUF - 1; ++r) issue_load(r);
  // Round r's first sweep runs BEFORE round r - 1's second sweep, and the two L2 round trips the second sweep needs
  // (the image's counter, then its slots) are started ahead of the work that can hide them: the counter is read
  // before the accumulation of round r, the slots before its block reduction.
  float acc[2][VE];
  float4 sl[GR_NSL];
  for (int r = 0; r <= p.nround; ++r) {
    GrPart pt, pp;
    pt.nrows = pp.nrows = 0;
    pt.img = pp.img = 0;
    pt.expected = pp.expected = 0;
    if (r < p.nround) pt = gr_part(p, gr_round(p, r), cta);
    if (r >= 1) pp = gr_part(p, gr_round(p, r - 1), cta);
    int seen = 0;
    GR_T(c0);
    // early look at the counter of round r - 1: a relaxed load (no wait here); normally every contributor has
    // published long ago.  The slot reads that follow are L2 loads issued after the branch on this value.
    if (pp.nrows > 0 && (tid & 31) == 0)
      asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(count + pp.img) : "memory");
    GR_T(c1);
    if (pt.nrows > 0) {
      const int b = r % GR_NBUF;
      st_mbar_wait(&full[b], (use_par >> b) & 1u);
      use_par ^= 1u << b;
    }
    GR_T(c2);
    if (pt.nrows > 0) accumulate(r, pt, acc);
    GR_T(c3);
    if (pp.nrows > 0) wait_published(pp, seen);
    GR_T(c4);
    if (pp.nrows > 0) fetch_slots(pp, sl);
    GR_T(c5);
    if (pt.nrows > 0) publish(pt, acc);
    GR_T(c6);
    if (pp.nrows > 0) totals(pp, sl);
    GR_T(c7);
    if (pp.nrows > 0) rescale(r - 1, pp);
    GR_T(c8);
    if (r >= 1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's in-place results -> copy engine
      __syncthreads();  // every thread is done with buffer (r - 1) % GR_NBUF
      if (tid == 0) {
        issue_store(r - 1);
        // all but the newest store group have finished READING shared memory: round r - 2's buffer is free
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        issue_load(r - 2 + GR_NBUF);
      }
    }
    GR_T(c9);
    GR_ADD(0, c0, c1); GR_ADD(1, c1, c2); GR_ADD(2, c2, c3); GR_ADD(3, c3, c4); GR_ADD(4, c4, c5);
    GR_ADD(5, c5, c6); GR_ADD(6, c6, c7); GR_ADD(7, c7, c8); GR_ADD(8, c8, c9); GR_ADD(9, c0, c0 + 1);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory outlives the stores
}

// Host side: rounds, buffer size and workspace layout.  false: the shape is not one for these kernels.
template <typename T, bool BWD>
bool grid_plan(int64_t B, int64_t C, int64_t L, int sms, GridPlan* plan, size_t* ws_bytes) {
  constexpr int VE = Vec16<T>::N;
  if (C % VE != 0 || B < 1 || L < 2 || B > GR_MAX_BATCH || L > (1 << 24) || sms < 8) return false;
  const int64_t cvec = C / VE;
  if (cvec != 8 && cvec != 16 && cvec != 32 && cvec != 64) return false;
  const int64_t row = C * (int64_t)sizeof(T), total = B * L * row;
  if (total < (8 << 20)) return false;  // small tensors: launch / ramp latency dominates, the cluster kernels do as well
  const int red = (GR_THREADS / 32) * 2 * VE * (int)(cvec < 32 ? cvec : 32) * 4;
  const int tail = red + 2 * (2 * (int)C * 4) + 64;
  const int64_t max_buf = ((GR_MAX_SMEM - tail) / GR_NBUF) / 256 * 256;
  const int64_t share_rows = max_buf / (row * (BWD ? 2 : 1));
  if (share_rows < 1) return false;
  // pieces per image m (one piece per CTA per round, k = sms / m images per round): the fullest machine over the
  // whole batch, minus 1 % per round (every round costs one grid-wide exchange of moments)
  const int64_t m_min = (L + share_rows - 1) / share_rows;
  double best = -1.0;
  int64_t best_m = 0;
  const int64_t groups = GR_THREADS / (2 * C / 4);  // slot groups of the totals pass (GrSmem::GROUPS)
  for (int64_t m = m_min; m <= sms && m <= L; ++m) {
    if ((m + groups - 1) / groups > GR_NSL) break;  // an image's slots are read GR_NSL per thread
    const int64_t k = sms / m < B ? sms / m : B, nround = (B + k - 1) / k;
    const double score = (double)(B * m) / (double)(nround * sms) - 0.01 * (double)nround;
    if (score > best + 1e-9) {
      best = score;
      best_m = m;
    }
  }
  if (best_m == 0) return false;
  const int64_t m = best_m, k = sms / m < B ? sms / m : B, rpi = (L + m - 1) / m;
  plan->B = (int)B;
  plan->L = (int)L;
  plan->nround = (int)((B + k - 1) / k);
  plan->m = (int)m;
  plan->rpi = (int)rpi;
  plan->buf_bytes = (int)((rpi * row * (BWD ? 2 : 1) + 255) / 256 * 256);
  if (plan->buf_bytes > max_buf) return false;
  *ws_bytes = (size_t)GR_COUNTER_BYTES + (size_t)B * m * 2 * C * sizeof(float);
  return true;
}

template <typename T, bool BWD, int CVEC>
int nlc_grid_launch(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out, const GridPlan& plan,
                    void* ws, int sms, float e_lambda, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  using S = GrSmem<VE, CVEC>;
  const int smem = GR_NBUF * plan.buf_bytes + S::TAIL_BYTES;
  auto kfun = &simam_nlc_grid<T, CVEC, BWD>;
  if (opt_in_smem(reinterpret_cast<const void*>(kfun), smem) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  // workspace: the per-image counters (zero between calls; always at the same place whatever the shape, because
  // the partial moments of one call must not be mistaken for counters by the next), then [B][m][2 C] floats
  int* count = static_cast<int*>(ws);
  int* done = count + GR_MAX_BATCH;
  float* gpart = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + GR_COUNTER_BYTES);  // [B][m][2 C]
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)sms);
  cfg.blockDim = dim3(GR_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  // Every CTA spins on its peers' counters, so all of them must become resident: one CTA per SM (shared memory)
  // and grid == #SMs guarantee it as soon as whatever else is running on the device drains.  A cooperative launch
  // would enforce it, but costs ~11 us per launch here (measured) — more than a third of the kernel.
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 0;
  cfg.attrs = at;
  cfg.numAttrs = getenv("CSB_SIMAM_COOP") ? 1 : 0;
  if (cfg.numAttrs) at[0].val.cooperative = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kfun, x, gy, stats_in, stats_out, out, plan, gpart, count, done, e_lambda);
  if (e != cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fail(CSB200_ERR_CUDA, "simam_nlc_grid: %s", cudaGetErrorString(e));
  }
  return check_launch(BWD ? "simam_nlc_grid<bwd>" : "simam_nlc_grid<fwd>");
}

// >= 0: launched (status code); -1: not applicable (shape, alignment, workspace) -> the caller falls through
template <typename T, bool BWD>
int nlc_grid(const T* x, const T* gy, float* stats_out, const float* stats_in, T* out, int64_t B, int64_t C, int64_t L,
             float e_lambda, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int VE = Vec16<T>::N;
  if (ws == nullptr || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(gy)) || !aligned16(ws)) return -1;
  const int sms = device_sm_count();
  GridPlan plan;
  size_t need = 0;
  if (sms <= 0 || !grid_plan<T, BWD>(B, C, L, sms, &plan, &need) || ws_bytes < need) return -1;
#define CSB_GRID(CV) \
  if (C / VE == CV) return nlc_grid_launch<T, BWD, CV>(x, gy, stats_out, stats_in, out, plan, ws, sms, e_lambda, st);
  CSB_GRID(8)
  CSB_GRID(16)
  CSB_GRID(32)
  CSB_GRID(64)
#undef CSB_GRID
  return -1;
}
