// Three-tile variant of K3 (round 1, measured +1 %: DESIGN.md section 4) -- evicted from the product library in round 2.
// It was a namespace inside lto_indirect_cw.cu and is kept here for the record; it does not build on its own.
// ---------------------------------------------------------------------------
// K3 v3 (experimental, LTO_K3=v3): THREE tiles in flight.  Per-warp cycle counters of the two-tile kernel above (tools/icw_prof.py) show the
// state warp's dependent chain as the bound: work 18.4 k + controller/refill 4.8 k cycles per attempt against 17.8 k
// cycles of column work per tile visit, so the column warps idle ~19 % of the time.  With three tiles a tile's next
// attempt has two column visits (35.6 k cycles) to be ready.  Still two state warps (8 warps, 255 registers): state
// warp 0 serves tiles 0 and 2 alternately, state warp 1 tile 1.  Three tiles of stage records leave no shared memory
// for the candidate stash: current and candidate columns live in two L2-resident buffers per (tile, half-phase,
// thread) and an accepted step flips which one is current (as in lto_indirect_cw14.cu).
// ---------------------------------------------------------------------------
__device__ __noinline__ Out9 sc_eval_state_call(double r0, double r1, double r2, double v0, double v1, double v2, double l0, double l1, double l2,
                                                double m0, double m1, double m2, double mu, double mu1, double omega, double pexp, double aL,
                                                double rho_inv, double rq);

namespace v3 {

constexpr int NT3 = 3;
constexpr size_t TILE3_BYTES = REC_BYTES + HDR_BYTES + ERR_BYTES + XN_BYTES;
constexpr size_t BAR3_BYTES = 64;
constexpr size_t STG_OFFSET = NT3 * TILE3_BYTES + NT3 * BAR3_BYTES;                  // cp.async staging: one column per column thread
constexpr size_t STG_BYTES = (size_t)ND * NCT * sizeof(double);
constexpr size_t SMEM3 = STG_OFFSET + STG_BYTES;
static_assert(SMEM3 <= 232448, "three tiles + staging must fit the 227 KB of shared memory per CTA");
static_assert(STG_OFFSET % 16 == 0, "staging area must be 16-byte aligned");
constexpr size_t SCRATCH3_DOUBLES_PER_CTA = (size_t)NT3 * 2 * 2 * ND * NCT;     // [tile][half][parity][component][thread]

__device__ __forceinline__ TileSmem tile_smem3(unsigned char* base, int t) {
    unsigned char* p = base + (size_t)t * TILE3_BYTES;
    TileSmem s;
    s.rec = reinterpret_cast<double2*>(p); p += REC_BYTES;
    s.hval = reinterpret_cast<double*>(p); p += TS * sizeof(double);
    s.hctl = reinterpret_cast<int2*>(p); p += TS * sizeof(int2);
    s.cur = nullptr; s.nx = nullptr;
    s.errp = reinterpret_cast<double*>(p); p += ERR_BYTES;
    s.xn = reinterpret_cast<double*>(p);
    unsigned char* b = base + (size_t)NT3 * TILE3_BYTES + (size_t)t * BAR3_BYTES;
    s.bar_full = smem_u32(b); s.bar_done = smem_u32(b + 8);
    s.tile_done = reinterpret_cast<volatile int*>(b + 16);
    return s;
}

// Column warps of the three-tile kernel.  A column's current value for the NEXT half-phase is fetched from its L2 buffer into a
// shared-memory staging area by cp.async while the present half-phase is being computed, so the L2 latency (~950 cycles per
// half-phase when loaded on demand: measured) is off the critical path.  Little state is kept across the attempt (the column
// arithmetic needs every register): tile pointers are recomputed from the tile index.
struct HalfJob { int2 hc; double h; };
constexpr int NP = ND / 2;                                              // double2 per column

__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool job_needs_column(const HalfJob& j) { return (j.hc.x & F_STORE) || ((j.hc.x & F_ACTIVE) && !(j.hc.x & F_RESET)); }
// buffer of (tile, half, which): which = 0 current, 1 candidate, given the parity word
__device__ __forceinline__ double2* col_buf(double2* scr, int t, int hf, unsigned par, int which) {
    const unsigned sel = ((par >> (2 * t + hf)) & 1u) ^ (unsigned)which;
    return scr + (size_t)(((t * 2 + hf) * 2) + sel) * NP * NCT;
}
// decode a half-phase's header (an accepted step flips the parity) and start fetching its current column
__device__ __forceinline__ HalfJob prefetch_half(unsigned char* smem, double2* scr, unsigned stg_u32, int t, int hf, int lane, unsigned& par) {
    const TileSmem S = tile_smem3(smem, t);
    const int slot = hf * HS + (lane & (HS - 1));
    HalfJob j;
    j.hc = S.hctl[slot];
    j.h = S.hval[slot];
    if (j.hc.x & F_ACCEPT) par ^= 1u << (2 * t + hf);
    if (job_needs_column(j)) {
        const double2* cur = col_buf(scr, t, hf, par, 0);
#pragma unroll
        for (int i = 0; i < NP; ++i)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg_u32 + (unsigned)(i * NCT * sizeof(double2))), "l"(cur + i * NCT) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return j;
}

template <bool JOINT>
__device__ __forceinline__ void column_warp3(const IndirectArgs& a, int cw, int lane, unsigned char* smem) {
    const int col = 2 * cw + (lane >> 4);
    const int ct = cw * 32 + lane;
    double2* const scr = reinterpret_cast<double2*>(a.scratch + (size_t)blockIdx.x * SCRATCH3_DOUBLES_PER_CTA) + ct;   // [tile][half][parity][NP][thread]
    const double2* const stg = reinterpret_cast<const double2*>(smem + STG_OFFSET) + ct;                           // [NP][thread]
    const unsigned stg_u32 = smem_u32(stg);
    unsigned alive = (1u << NT3) - 1u;
    unsigned visit = 0;
    unsigned par = 0;                                                   // bit (2 t + hf): which buffer holds the current column
    long long c_wait = 0, n_work = 0;
    const long long c_begin = clock64();
    int t = 0, hf = 0;
    { const long long c0 = clock64(); mbar_wait_parked(tile_smem3(smem, t).bar_full, visit & 1); c_wait += clock64() - c0; }
    bool done = *tile_smem3(smem, t).tile_done != 0;
    HalfJob j = prefetch_half(smem, scr, stg_u32, t, 0, lane, par);
    // ONE loop body per half-phase (a single copy of the ~30 KB column code: the instruction cache matters here)
    while (true) {
        // ---- this half-phase's column: out of the staging area, which is then free for the next fetch
        const unsigned par_now = par;
        double p[ND];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (job_needs_column(j)) {
#pragma unroll
            for (int i = 0; i < NP; ++i) { const double2 v = stg[i * NCT]; p[2 * i] = v.x; p[2 * i + 1] = v.y; }
        } else {
#pragma unroll
            for (int i = 0; i < ND; ++i) p[i] = 0.0;
        }
        // ---- what comes next, and its fetch (overlaps the arithmetic below)
        int tn = t; unsigned vn = visit;
        bool have_next = false, done_n = done;
        HalfJob jn = j;
        if (hf == 0) {
            jn = prefetch_half(smem, scr, stg_u32, t, 1, lane, par);
            have_next = true;
        } else {
            const unsigned rest = alive & ~(1u << t);
            const unsigned higher = rest & ~((2u << t) - 1u);
            if (higher) tn = __ffs(higher) - 1;
            else { vn = visit + 1; const unsigned wrap = done ? rest : alive; tn = wrap ? __ffs(wrap) - 1 : -1; }
            if (tn >= 0 && tn != t) {
                const TileSmem Sn = tile_smem3(smem, tn);
                if (mbar_test(Sn.bar_full, vn & 1)) {                  // usually already published: two column visits of slack
                    done_n = *Sn.tile_done != 0;
                    jn = prefetch_half(smem, scr, stg_u32, tn, 0, lane, par);
                    have_next = true;
                }
            }
        }
        // ---- this half-phase
        {
            const TileSmem S = tile_smem3(smem, t);
            const int slot = hf * HS + (lane & (HS - 1));
            if (j.hc.x & F_STORE) {                                    // column `col` of ForwardDiff.jacobian(f, x0) (:121)
                double* out = a.phi + (long long)j.hc.y * (ND * ND) + col * ND;
                if ((reinterpret_cast<uintptr_t>(a.phi) & 31u) == 0) {
#pragma unroll
                    for (int i = 0; i < ND; i += 4)
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out + i), "d"(p[i]), "d"(p[i + 1]), "d"(p[i + 2]), "d"(p[i + 3]) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < ND; ++i) out[i] = p[i];
                }
            }
            if (j.hc.x & F_RESET) {
                double2* cur = col_buf(scr, t, hf, par_now, 0);
#pragma unroll
                for (int i = 0; i < ND; ++i) p[i] = (i == col) ? 1.0 : 0.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) __stcg(cur + i * NCT, make_double2(p[2 * i], p[2 * i + 1]));
            }
            if (!done) {
                double pn[ND];
                const double es = col_attempt<JOINT>(p, j.h, 2.0 * a.c.omega, S.rec + slot, S.bar_full, 0u, a.cfg.atol, a.cfg.rtol, pn);
                double2* cand = col_buf(scr, t, hf, par_now, 1);
#pragma unroll
                for (int i = 0; i < NP; ++i) __stcg(cand + i * NCT, make_double2(pn[2 * i], pn[2 * i + 1]));
                if (JOINT) S.errp[col * TS + slot] = es;
                ++n_work;
            }
        }
        // ---- advance
        if (hf == 0) { hf = 1; j = jn; continue; }
        if (done) alive &= ~(1u << t);
        else mbar_arrive(tile_smem3(smem, t).bar_done);
        if (tn < 0 || !alive) break;
        if (!have_next) {
            const TileSmem Sn = tile_smem3(smem, tn);
            const long long c0 = clock64();
            mbar_wait_parked(Sn.bar_full, vn & 1);
            c_wait += clock64() - c0;
            done_n = *Sn.tile_done != 0;
            jn = prefetch_half(smem, scr, stg_u32, tn, 0, lane, par);
        }
        t = tn; visit = vn; done = done_n; hf = 0; j = jn;
    }
    if (a.prof && lane == 0) {
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NW + NTILE + cw) * 4;
        const long long tot = clock64() - c_begin;
        o[0] = tot - c_wait; o[1] = c_wait; o[2] = n_work; o[3] = tot;
        a.prof[(size_t)gridDim.x * NW * 4 + (size_t)gridDim.x * NTILE + (size_t)blockIdx.x * NCW + cw] = 0;
    }
}

struct SlotCtl {
    double tcur, tf, h, span, esum, aL, rho_inv, rq;
    long long seg, ia;
    int na, nt, status, xi;
    bool active, lastrej, last, have;
    unsigned visit;
};

// state warp `sw` serves tiles sw, sw + 2, ... (warp 0: tiles 0 and 2; warp 1: tile 1)
template <bool JOINT>
__device__ __forceinline__ void state_warp3(const IndirectArgs& a, int sw, int lane, unsigned char* smem) {
    const int slot = lane;
    const unsigned fullmask = 0xffffffffu;
    const double atol = a.cfg.atol, rtol = a.cfg.rtol;
    const double inv_ne = JOINT ? 1.0 / (double)(ND * (ND + 1)) : 1.0 / (double)ND;
    SlotCtl ctl[2];
    unsigned alive = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        SlotCtl& c = ctl[k];
        c.tcur = 0.0; c.tf = 0.0; c.h = 0.0; c.span = 1.0; c.esum = 0.0; c.aL = 0.0; c.rho_inv = 1.0; c.rq = 0.0;
        c.seg = -1; c.ia = 0; c.na = 0; c.nt = 0; c.status = 0; c.xi = 0;
        c.active = false; c.lastrej = false; c.last = false; c.have = false; c.visit = 0;
        const int t = sw + 2 * k;
        if (t < NT3) {
            alive |= 1u << k;
            double* const xb = tile_smem3(smem, t).xn + slot;
#pragma unroll
            for (int i = 0; i < ND; ++i) { xb[i * TS] = 0.0; xb[(ND + i) * TS] = 0.0; }
        }
    }
    bool exhausted = false;
    long long c_wait = 0, c_work = 0, c_pre = 0, n_att = 0;
    const long long c_begin = clock64();
    while (alive) {
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            if (!(alive & (1u << k))) continue;
            const TileSmem S = tile_smem3(smem, sw + 2 * k);
            SlotCtl c = ctl[k];
            double* const xbuf = S.xn + slot;
            double2* const rec = S.rec + slot;
            int flags = 0, store_seg = 0;
            bool finished = false;
            const long long c0 = clock64();
            long long c1 = c0;
            if (c.have) {
                mbar_wait_parked(S.bar_done, (c.visit - 1) & 1);
                c1 = clock64();
                c_wait += c1 - c0;
                if (c.active) {
                    double s2 = c.esum;
                    if (JOINT) {
#pragma unroll
                        for (int j = 0; j < ND; ++j) s2 += S.errp[j * TS + slot];
                    }
                    const double eest = sqrt(s2 * inv_ne);
                    if (!(eest == eest)) { c.status = LTO_ST_NAN; finished = true; }
                    else {
                        double q = (eest == 0.0) ? 5.0 : 0.9 * inv_eighth_root(eest);
                        q = fmin(5.0, fmax(0.2, q));
                        if (eest <= 1.0) {
                            ++c.na; flags |= F_ACCEPT;
                            c.xi ^= 1;
                            if (c.last) { c.tcur = c.tf; finished = true; }
                            else { c.tcur += c.h; if (c.lastrej) q = fmin(q, 1.0); c.lastrej = false; }
                        } else {
                            c.lastrej = true; q = fmin(q, 1.0);
                        }
                        c.h *= q;
                    }
                }
            }
            if (c.active && !finished) {                                   // drive_rk8's loop-top checks
                if (c.h < c.span * 1e-12) { c.status = LTO_ST_HMIN; finished = true; }
                else if (c.nt >= a.cfg.max_attempts) { c.status = LTO_ST_MAXSTEPS; finished = true; }
            }
            if (c.active && finished) {
                // ---- defect = x(t1) - XC_all[:, i+1] (multiShoot_CRTBP_indirect.jl:82)
                bool nan = false;
                const double* xs = xbuf + c.xi * ND * TS;
#pragma unroll
                for (int i = 0; i < ND; ++i) {
                    const double xv = xs[i * TS];
                    nan |= !(xv == xv);
                    a.defect[c.seg * ND + i] = a.x_target ? xv - a.x_target[c.ia * ND + i] : xv;
                }
                if (nan && c.status == 0) c.status = LTO_ST_NAN;
                if (a.status) a.status[c.seg] = c.status;
                if (a.nsteps_out) { a.nsteps_out[2 * c.seg] = c.na; a.nsteps_out[2 * c.seg + 1] = c.nt; }
                flags |= F_STORE; store_seg = (int)c.seg;
                c.active = false;
            }
            bool fresh = false;
            if (!c.active && !exhausted) {
                const long long idx = (long long)atomicAdd(a.counter, 1ull);
                if (idx < a.n_seg) {
                    c.seg = idx; c.ia = lto_node_a(c.seg, a.npt);
                    const long long it = lto_traj_of(c.seg, a.npt);
#pragma unroll
                    for (int i = 0; i < ND; ++i) xbuf[(c.xi * ND + i) * TS] = a.x0[c.ia * ND + i];
                    c.tcur = a.t0[c.ia]; c.tf = a.t1[c.ia];
                    if (!(c.tcur < c.tf)) c.tf = c.tcur;                    // empty span: one zero-length step, Phi = I
                    c.span = c.tf - c.tcur;
                    const double tl = a.thrustLimit_arr ? a.thrustLimit_arr[it] : a.c.thrustLimit;
                    const double rho = a.rho_arr ? a.rho_arr[it] : a.c.rho;
                    c.aL = tl * a.c.kthr / a.c.mass;                        // :33
                    c.rho_inv = 1.0 / rho;
                    c.rq = c.aL / (4.0 * rho);
                    c.na = 0; c.nt = 0; c.status = 0; c.lastrej = false;
                    c.active = true; fresh = true; flags |= F_RESET;
                } else {
                    exhausted = true;
                }
            }
            if (!__any_sync(fullmask, c.active)) {
                S.hval[slot] = 0.0; S.hctl[slot] = make_int2(flags, store_seg);
                if (lane == 0) *S.tile_done = 1;
                mbar_arrive(S.bar_full);
                alive &= ~(1u << k);
                ctl[k] = c;
                continue;
            }
            LawConst lw; lw.aL = c.aL; lw.rho_inv = c.rho_inv; lw.rho_inv_quarter_aL = c.rq;
            KStore K;
            const double* xs = xbuf + c.xi * ND * TS;
            const long long c2 = clock64();
            c_pre += c2 - c1;
            state_stage_s<0>(K, xs, 0.0, 0.0, a.c, lw, rec, S.bar_full);
            if (__any_sync(fullmask, fresh)) {
                // Hairer-Norsett-Wanner initial step over the state components (drive_rk8 in lto_prop_generic.cuh)
                double f0[ND], y1[ND], x[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = xs[i * TS];
#pragma unroll
                for (int q = 0; q < 3; ++q) { f0[q] = x[3 + q]; f0[3 + q] = K.kv[0][q]; f0[6 + q] = K.kl[0][q]; f0[9 + q] = K.km[0][q]; }
                const double d0 = rms12(x, x, atol, rtol), d1 = rms12(f0, x, atol, rtol);
                double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                h0 = fmin(h0, c.span);
#pragma unroll
                for (int i = 0; i < ND; ++i) y1[i] = fma(h0, f0[i], x[i]);
                {
                    const Out9 o = sc_eval_state_call(y1[0], y1[1], y1[2], y1[3], y1[4], y1[5], y1[6], y1[7], y1[8], y1[9], y1[10], y1[11], a.c.mu, a.c.m1,
                                                      a.c.omega, a.c.p, lw.aL, lw.rho_inv, lw.rho_inv_quarter_aL);
#pragma unroll
                    for (int q = 0; q < 3; ++q) { const double v1 = y1[3 + q]; y1[q] = v1 - f0[q]; y1[3 + q] = o.v[q] - f0[3 + q]; y1[6 + q] = o.v[3 + q] - f0[6 + q]; y1[9 + q] = o.v[6 + q] - f0[9 + q]; }
                }
                const double d2 = rms12(y1, x, atol, rtol) / h0;
                const double dm = fmax(d1, d2);
                const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : inv_eighth_root(dm / 0.01);
                if (fresh) c.h = fmin(fmin(100.0 * h0, h1), c.span);
            }
            c.last = false;
            if (c.tcur + c.h >= c.tf) { c.h = c.tf - c.tcur; c.last = true; }
            if (c.active) ++c.nt;
            const double h = c.h;
            S.hval[slot] = h; S.hctl[slot] = make_int2(flags | (c.active ? F_ACTIVE : 0), store_seg);
            const double h2 = h * h;
            state_stage_s<1>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<2>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<3>(K, xs, h, h2, a.c, lw, rec, S.bar_full);
            state_stage_s<4>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<5>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<6>(K, xs, h, h2, a.c, lw, rec, S.bar_full);
            state_stage_s<7>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<8>(K, xs, h, h2, a.c, lw, rec, S.bar_full);  state_stage_s<9>(K, xs, h, h2, a.c, lw, rec, S.bar_full);
            state_stage_s<10>(K, xs, h, h2, a.c, lw, rec, S.bar_full); state_stage_s<11>(K, xs, h, h2, a.c, lw, rec, S.bar_full); state_stage_s<12>(K, xs, h, h2, a.c, lw, rec, S.bar_full);
            {
                double x[ND], xn[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) x[i] = xs[i * TS];
                c.esum = step_finish<true>(K, x, h, h2, atol, rtol, xn);
                double* xc = xbuf + (c.xi ^ 1) * ND * TS;
#pragma unroll
                for (int i = 0; i < ND; ++i) xc[i * TS] = xn[i];
            }
            mbar_arrive(S.bar_full);
            c_work += clock64() - c2; ++n_att;
            c.have = true; ++c.visit;
            ctl[k] = c;
        }
    }
    if (a.prof && lane == 0) {
        unsigned long long* o = a.prof + ((size_t)blockIdx.x * NW + sw) * 4;
        o[0] = c_work; o[1] = c_wait; o[2] = n_att; o[3] = clock64() - c_begin;
        a.prof[(size_t)gridDim.x * NW * 4 + (size_t)blockIdx.x * NTILE + sw] = c_pre;
    }
}

template <bool JOINT>
__global__ void __launch_bounds__(NTHREADS, 1) k_indirect_cw3(IndirectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NT3) {
        const TileSmem S = tile_smem3(smem_raw, threadIdx.x);
        mbar_init(S.bar_full, 32);
        mbar_init(S.bar_done, NCT);
        *S.tile_done = 0;
    }
    __syncthreads();
    // warps 3 and 7 (both on SM sub-partition 3) are the state warps, as in the two-tile kernel
    if ((warp & 3) == 3) state_warp3<JOINT>(a, warp >> 2, lane, smem_raw);
    else column_warp3<JOINT>(a, warp - (warp >> 2), lane, smem_raw);
}

}  // namespace v3

