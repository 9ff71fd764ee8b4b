// Lane re-packing photon kernel (included by mmcb_kernel.cu after the shared helpers).
//
// The flattened kernel (mmcb_photon_kernel) runs step, hop, scatter and deposit flush under divergence: ncu counted 22.9 of 32
// active lanes per warp instruction on the headline workload and ~445 warp instructions per iteration.  Here a warp owns 64
// WALKERS (photon + its xorshift128+ stream: what a thread of the reference is, src/mmc_core.cl:2163-2215) for its 32 lanes:
// 32 sit in registers, 32 in a per-warp shared-memory stash.  At the end of every step a walker's NEXT step is classified from
// the element record:
//     F  it ends inside the element at a scattering site, within the time window, and its deposit goes to ONE accumulator
//        (element mode: always; dual grid: both segment midpoints in one voxel): move, attenuate, add to the pending run
//        (closing it first when the accumulator changed), scatter -- no face search, no division, no segment loop, no hop;
//     G  anything else (face crossing, reflection, steps that straddle voxels, end of the time window, launch): the general step.
// Every round the warp picks the class that can fill its lanes (|F| + |G| = 64, so one of them has >= 32 walkers), swaps the
// lanes holding the other class with stash slots of the chosen one (five 128-bit shared-memory exchanges, no block barrier) and
// runs only that class's code, converged.  Walkers never share a stream, and a walker uses its stream exactly like a reference
// thread does, so the RNG sequence per walker stays bit-exact.
#pragma once

#define MMCB_RP_SLOTS 32                // stash slots per warp (= walkers per warp - 32)
#ifndef MMCB_RP_THREADS
#define MMCB_RP_THREADS 256
#endif
#ifndef MMCB_RP_MINBLOCKS
#define MMCB_RP_MINBLOCKS 4
#endif
#ifndef MMCB_RP_MINBLOCKS_DET
#define MMCB_RP_MINBLOCKS_DET 3
#endif

// walker classes
#define RP_F     0      // in flight, next step is a fast step
#define RP_G     1      // in flight, next step needs the general code
#define RP_NEED  2      // needs a photon (general round)
#define RP_DEAD  3      // no photons left

struct RpExtra {            // per-walker state beyond Photon + Rng
    int mode, type;
    int acct;               // DET: medium of the running partial-path sums
    float accL, accN, accM;
    Rng initseed;           // DET: stream state at launch (saved with the detected photon)
    int col;                // DET: the walker's column of the shared-memory partial-path table
    unsigned int nidx;      // class F: accumulator index of the coming step (classification result)
};

// swap four registers with a 16-byte shared-memory word (shared-space address + immediate offset: no address arithmetic per chunk)
#define RP_SWAP4(saddr, off, a0, a1, a2, a3) do { \
        unsigned int t0_, t1_, t2_, t3_; \
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+" #off "];" : "=r"(t0_), "=r"(t1_), "=r"(t2_), "=r"(t3_) : "r"(saddr)); \
        asm volatile("st.shared.v4.u32 [%0+" #off "], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory"); \
        a0 = t0_; a1 = t1_; a2 = t2_; a3 = t3_; \
    } while (0)

template <bool DET>
__device__ __forceinline__ void rp_exchange(unsigned int q, Photon& p, Rng& rng, RpExtra& x) {
    // swap the register walker with the stash slot at shared-space address q (chunks are MMCB_RP_SLOTS * 16 = 512 bytes apart:
    // conflict-free 128-bit accesses)
    unsigned int a0, a1, a2, a3;
    a0 = __float_as_uint(p.px), a1 = __float_as_uint(p.py), a2 = __float_as_uint(p.pz), a3 = __float_as_uint(p.vx);
    RP_SWAP4(q, 0, a0, a1, a2, a3);
    p.px = __uint_as_float(a0), p.py = __uint_as_float(a1), p.pz = __uint_as_float(a2), p.vx = __uint_as_float(a3);
    a0 = __float_as_uint(p.vy), a1 = __float_as_uint(p.vz), a2 = __float_as_uint(p.w), a3 = __float_as_uint(p.t);
    RP_SWAP4(q, 512, a0, a1, a2, a3);
    p.vy = __uint_as_float(a0), p.vz = __uint_as_float(a1), p.w = __uint_as_float(a2), p.t = __uint_as_float(a3);
    a0 = __float_as_uint(p.slen), a1 = (unsigned int)p.eid, a2 = p.oldidx, a3 = __float_as_uint(p.oldw);
    RP_SWAP4(q, 1024, a0, a1, a2, a3);
    p.slen = __uint_as_float(a0), p.eid = (int)a1, p.oldidx = a2, p.oldw = __uint_as_float(a3);
    a0 = (unsigned int)(rng.t0 >> 32), a1 = (unsigned int)rng.t0, a2 = (unsigned int)(rng.t1 >> 32), a3 = (unsigned int)rng.t1;
    RP_SWAP4(q, 1536, a0, a1, a2, a3);
    rng.t0 = ((unsigned long long)a0 << 32) | a1;
    rng.t1 = ((unsigned long long)a2 << 32) | a3;
    a0 = (unsigned int)p.fixcount, a1 = (unsigned int)x.type, a2 = (unsigned int)x.mode | ((unsigned int)x.col << 2), a3 = x.nidx;
    RP_SWAP4(q, 2048, a0, a1, a2, a3);
    p.fixcount = (int)a0, x.type = (int)a1, x.mode = (int)(a2 & 3u), x.col = (int)(a2 >> 2), x.nidx = a3;

    if (DET) {
        a0 = (unsigned int)x.acct, a1 = __float_as_uint(x.accL), a2 = __float_as_uint(x.accN), a3 = __float_as_uint(x.accM);
        RP_SWAP4(q, 2560, a0, a1, a2, a3);
        x.acct = (int)a0, x.accL = __uint_as_float(a1), x.accN = __uint_as_float(a2), x.accM = __uint_as_float(a3);
        a0 = (unsigned int)(x.initseed.t0 >> 32), a1 = (unsigned int)x.initseed.t0, a2 = (unsigned int)(x.initseed.t1 >> 32), a3 = (unsigned int)x.initseed.t1;
        RP_SWAP4(q, 3072, a0, a1, a2, a3);
        x.initseed.t0 = ((unsigned long long)a0 << 32) | a1;
        x.initseed.t1 = ((unsigned long long)a2 << 32) | a3;
    }
}

// METHOD: 3 branch-less Badouel with per-element deposit, 4 BLB with dual-grid (DMMC) deposit.  Single-pattern sources of every
// type, no replay / trajectories / diffuse reflectance / multi-slot / RF (those keep mmcb_photon_kernel<.., GENERAL = true>).
template <int METHOD, bool DET>
__global__ void __launch_bounds__(MMCB_RP_THREADS, DET ? MMCB_RP_MINBLOCKS_DET : MMCB_RP_MINBLOCKS)
mmcb_photon_kernel_rp(const mmcb_kargs a) {
    constexpr bool GRID = (METHOD == 4);
    constexpr int NCH = DET ? 7 : 5;                // 128-bit chunks per stashed walker
    const bool hoton = gp.hotcache != 0 && a.hotstat[MMCB_HOT_STAT_USEFUL] != 0;
    const uint2 hot = hoton ? make_uint2(a.hotstat[MMCB_HOT_STAT_LO], a.hotstat[MMCB_HOT_STAT_SPAN]) : make_uint2(0u, 0u);
    unsigned int* hkeys = (unsigned int*)smem4;
    float* hvals = (float*)smem4 + MMCB_HOT_SLOTS;
    float4* smed = smem4 + (gp.hotcache ? MMCB_HOT_BYTES / 16 : 0);     // media table, 2 float4 per medium
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* stash = (uint4*)(smed + 2 * gp.nmedia) + (size_t)warp * (NCH * MMCB_RP_SLOTS);      // this warp's stash: [chunk][slot]
    unsigned int* scratch = (unsigned int*)((uint4*)(smed + 2 * gp.nmedia) + (size_t)nwarp * (NCH * MMCB_RP_SLOTS)) + warp * 32;
    float* ppath = (float*)((unsigned int*)((uint4*)(smed + 2 * gp.nmedia) + (size_t)nwarp * (NCH * MMCB_RP_SLOTS)) + nwarp * 32);    // DET: [reclen][2 * blockDim]
    const unsigned int stash_s = (unsigned int)__cvta_generic_to_shared(stash);     // shared-space address of the warp's stash
    const int ncol = 2 * blockDim.x;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned ltmask = (1u << lane) - 1u;
    acc_t* field = (acc_t*)a.field;
    const unsigned long long gfield = (unsigned long long)__cvta_generic_to_global(a.field);

    for (int i = threadIdx.x; i < 2 * gp.nmedia; i += blockDim.x) {
        smed[i] = a.med[i];
    }

    if (hoton) {
        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS; i += blockDim.x) {
            hkeys[i] = a.hotkeys[i];
        }

        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS * MMCB_HOT_GROUP; i += blockDim.x) {
            hvals[i] = 0.f;
        }
    }

    // walker w of global warp gw: stream 64 gw + w; w < 32 starts in lane w, the others in stash slot w - 32
    const size_t gw = (size_t)blockIdx.x * nwarp + warp;
    Photon p;
    Rng rng;
    RpExtra x;
    p.px = p.py = p.pz = p.vx = p.vy = p.vz = p.w = p.t = p.slen = p.slen0 = p.oldw = 0.f;
    p.eid = 0;
    p.oldidx = 0xFFFFFFFFu;
    p.posidx = 0;
    p.id = 0;
    p.fixcount = 0;
    p.slotoff = 0;
    p.w_im = p.oldw_im = 0.f;
    x.mode = RP_NEED;
    x.type = 0;
    x.acct = 0;
    x.accL = x.accN = x.accM = 0.f;
    x.nidx = 0xFFFFFFFFu;
    x.col = warp * 64 + 32 + lane;
    {
        const uint4 s = *(const uint4*)(a.seeds + 4 * (gw * 64 + 32 + lane));
        rng.t0 = ((unsigned long long)s.x << 32) | s.y;
        rng.t1 = ((unsigned long long)s.z << 32) | s.w;
        x.initseed = rng;
        // park the second walker: the exchange writes the register walker and returns what the slot held (garbage, overwritten below)
        uint4* q = stash + lane;
        q[0] = make_uint4(0u, 0u, 0u, 0u);
        q[MMCB_RP_SLOTS] = make_uint4(0u, 0u, 0u, 0u);
        q[2 * MMCB_RP_SLOTS] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0u);
        q[3 * MMCB_RP_SLOTS] = s;
        q[4 * MMCB_RP_SLOTS] = make_uint4(0u, 0u, (unsigned int)RP_NEED | ((unsigned int)x.col << 2), 0u);

        if (DET) {
            q[5 * MMCB_RP_SLOTS] = make_uint4(0u, 0u, 0u, 0u);
            q[6 * MMCB_RP_SLOTS] = s;
        }

        const uint4 s2 = *(const uint4*)(a.seeds + 4 * (gw * 64 + lane));   // xorshift128p_seed, src/mmc_core.cl:545-548
        rng.t0 = ((unsigned long long)s2.x << 32) | s2.y;
        rng.t1 = ((unsigned long long)s2.z << 32) | s2.w;
        x.initseed = rng;
        x.col = warp * 64 + lane;
    }
    __syncthreads();

    unsigned int sF = 0u, sG = FULL;    // stash slots holding class-F walkers / class-G (RP_G or RP_NEED) walkers
    float etot = 0.f, eesc = 0.f;       // per-lane tallies like src/mmc_core.cl:1908,2155
    unsigned int nraytet = 0;
    unsigned int pool_next = 0, pool_end = 0;       // lane 0: the warp's range of photon ids (work stealing)
    const unsigned int nlaunch = (unsigned int)gp.nphoton;
    const int reclen = gp.reclen;
    const int M = gp.maxmedia;
#define PPATH(k) ppath[(k) * ncol + x.col]
#define PPATH_FLUSH() do { if (x.acct > 0 && x.acct <= M) { PPATH(M + x.acct - 1) += x.accL; PPATH(x.acct - 1) += x.accN; if (gp.ismomentum) { PPATH(2 * M + x.acct - 1) += x.accM; } } x.accL = x.accN = x.accM = 0.f; } while (0)

    while (true) {
        // ------------------------------------------------------------------ pick the round's class and fill the lanes with it
        const unsigned rF = __ballot_sync(FULL, x.mode == RP_F), rG = __ballot_sync(FULL, x.mode == RP_G || x.mode == RP_NEED);
        const int nF = __popc(rF) + __popc(sF), nG = __popc(rG) + __popc(sG);

        if (nF + nG == 0) {
            break;
        }

        const bool roundF = (nF >= nG);
        {
            const unsigned inmask = roundF ? sF : sG;
            const bool wrong = roundF ? (x.mode != RP_F) : !(x.mode == RP_G || x.mode == RP_NEED);
            const unsigned outb = roundF ? ~rF : ~rG;

            if (outb != 0u && inmask != 0u) {
                const int k = min(__popc(outb), __popc(inmask));

                if ((inmask >> lane) & 1u) {
                    scratch[__popc(inmask & ltmask)] = (unsigned int)lane;      // r-th stash slot of the wanted class
                }

                __syncwarp();
                const int r = __popc(outb & ltmask);
                const bool part = wrong && r < k;
                const int oldmode = x.mode;
                unsigned int bit = 0u;

                if (part) {
                    const unsigned int slot = scratch[r];
                    bit = 1u << slot;
                    rp_exchange<DET>(stash_s + slot * 16u, p, rng, x);
                }

                __syncwarp();
                // an F round sends class-G (or dead) walkers to the stash, a G round class-F (or dead) ones
                const unsigned taken = __reduce_or_sync(FULL, bit);
                const unsigned live = __reduce_or_sync(FULL, (oldmode != RP_DEAD) ? bit : 0u);
                sF = roundF ? (sF & ~taken) : (sF | live);
                sG = roundF ? (sG | live) : (sG & ~taken);
            }
        }

        bool classify = false;      // this lane's walker took a step this round and is still in flight

        if (roundF) {
            // -------------------------------------------------------------- fast step: ends inside the element at a scattering site
            if (x.mode == RP_F) {
                const float4 prop = smed[2 * x.type];           // mua mus g n
                const float4 pd = smed[2 * x.type + 1];         // 1/mus, n/c0, 1/mua (0: mua < EPS), c0/n
                const float Lmove = p.slen * pd.x;
                nraytet++;
                p.t += Lmove * pd.y;
                const float w0 = p.w;
                p.w *= __expf(-prop.x * Lmove);
                float ww = w0 - p.w;

                if (gp.outputtype != 2) {                       // src/mmc_core.cl:844-851
                    ww = (pd.z == 0.f) ? (w0 * Lmove) : (ww * pd.z);
                }

                if (x.nidx != p.oldidx) {                       // the step starts a new run (other voxel or gate, same-voxel midpoints)
                    if (p.oldw > 0.f) {
                        flush_deposit<false>(gfield, p.oldidx, p.oldw, p, a, hot);
                    }

                    p.oldidx = x.nidx;
                    p.oldw = 0.f;
                }

                p.oldw += ww;
                p.px += Lmove * p.vx;
                p.py += Lmove * p.vy;
                p.pz += Lmove * p.vz;
                p.fixcount = 0;

                if (DET) {
                    if (x.type != x.acct) {
                        PPATH_FLUSH();
                        x.acct = x.type;
                    }

                    x.accL += Lmove;
                }

                bool dead = false;

                if (p.w < gp.roulette_w) {                      // :2101-2114
                    if (rand01(rng) * gp.roulettesize <= 1.f) {
                        p.w *= gp.roulettesize;
                    } else {
                        dead = true;
                    }
                }

                if (dead) {
                    if (DET) {
                        PPATH_FLUSH();
                        x.acct = 0;
                    }

                    eesc += p.w;
                    x.mode = RP_NEED;       // the pending run is dropped like the reference does (it flushes on the next step, which never comes)
                } else {
                    float mom;
                    p.slen0 = next_scatter(prop.z, p, rng, mom);
                    p.slen = p.slen0;

                    if (DET) {
                        x.accM += mom;
                        x.accN += 1.f;
                    }

                    classify = true;
                }
            }
        } else {
            // -------------------------------------------------------------- general round: photon supply, then one full step
            const unsigned need = __ballot_sync(FULL, x.mode == RP_NEED);

            if (need) {
                unsigned int myid = 0;
                bool got = false;
                const int n = __popc(need);
                unsigned int base = 0;
                int avail = 0;

                if (lane == 0 && pool_end - pool_next < (unsigned int)n) {
                    base = pool_next;
                    avail = (int)(pool_end - pool_next);
                    const unsigned int want = (unsigned int)(POOL_CHUNK + n - avail);
                    const unsigned long long g0 = atom_add_u64(a.photon_counter, want);
                    pool_next = (unsigned int)min(g0, (unsigned long long)nlaunch);
                    pool_end = max((unsigned int)min(g0 + want, (unsigned long long)nlaunch), pool_next);
                }

                avail = __shfl_sync(FULL, avail, 0);
                base = __shfl_sync(FULL, base, 0);
                const unsigned int pn = __shfl_sync(FULL, pool_next, 0);
                const unsigned int pe = __shfl_sync(FULL, pool_end, 0);
                const int rank = __popc(need & ltmask);

                if (x.mode == RP_NEED) {
                    if (rank < avail) {
                        myid = base + rank;
                        got = true;
                    } else {
                        const unsigned int cand = pn + (unsigned int)(rank - avail);

                        if (cand < pe) {
                            myid = cand;
                            got = true;
                        }
                    }
                }

                if (lane == 0) {
                    pool_next = min(pool_next + (unsigned int)max(0, n - avail), pool_end);
                }

                if (x.mode == RP_NEED) {
                    if (got) {
                        p.id = myid + (unsigned int)gp.photon_offset;

                        if (DET) {
                            x.initseed = rng;

                            for (int k = 0; k < reclen; k++) {
                                PPATH(k) = 0.f;
                            }
                        }

                        launch_photon<false>(p, rng, a);

                        if (DET) {
                            PPATH(reclen - 1) = p.w;            // :1894-1898
                        }

                        etot += p.w;
                        x.mode = RP_G;
                    } else {
                        x.mode = RP_DEAD;
                    }
                }
            }

            int detid = 0;

            if (x.mode == RP_G) {
                nraytet++;
                float Lmove = 0.f;
                float4 prop = make_float4(0.f, 0.f, 0.f, 1.f);
                int neweid = 0, type = 0, faceidx = 0;
                unsigned flags = 0;
                bool found = false, isend = false, timeup = false;
                bool terminate = false, detect = false;
                {
                    const mmcb_tetrec* rec = a.tet + (p.eid - 1);
                    float r0[8], r1[8], r2[8];
                    ld256(rec, r0);                                     // nx[4] ny[4]
                    ld256((const char*)rec + 32, r1);                   // nz[4] d[4]
                    ld256((const char*)rec + 64, r2);                   // nb[4] type flags
                    float T[4];         // src/mmc_core.cl:752-771
                    #pragma unroll

                    for (int j = 0; j < 4; j++) {
                        const float S = p.vx * r0[j] + p.vy * r0[4 + j] + p.vz * r1[j];
                        const float Tn = r1[4 + j] - (p.px * r0[j] + p.py * r0[4 + j] + p.pz * r1[j]);
                        T[j] = (S > 0.f) ? __fdividef(Tn, S) : 1e10f;
                    }

                    const float Lmin = fminf(fminf(T[0], T[1]), fminf(T[2], T[3]));
                    faceidx = (T[0] == Lmin) ? 0 : ((T[1] == Lmin) ? 1 : ((T[2] == Lmin) ? 2 : 3));
                    neweid = __float_as_int((faceidx == 0) ? r2[0] : ((faceidx == 1) ? r2[1] : ((faceidx == 2) ? r2[2] : r2[3])));
                    type = __float_as_int(r2[4]);
                    flags = __float_as_uint(r2[5]) >> faceidx;      // bit 0: reflect, bit 4: to void, bit 8: from void
                    found = (Lmin < 1e10f && Lmin >= 0.f);

                    if (found) {
                        prop = smed[2 * type];
                        const float4 pd = smed[2 * type + 1];
                        Lmove = (pd.x == 0.f) ? R_MIN_MUS : p.slen * pd.x;
                        isend = (Lmin > Lmove);
                        Lmove = isend ? Lmove : Lmin;
                        const float rc = pd.y;
                        float tnew = p.t + Lmove * rc;
                        int gate = (int)((tnew - gp.tstart) * gp.Rtstep);

                        if (gate > gp.maxgate - 1) {                    // :803-807
                            timeup = true;
                            Lmove = (gp.tend - p.t) * pd.w - 1e-4f;
                            tnew = p.t + Lmove * rc;
                            gate = min((int)((tnew - gp.tstart) * gp.Rtstep), gp.maxgate - 1);
                        }

                        const float currweight = p.w;
                        float totalloss = __expf(-prop.x * Lmove);
                        p.w *= totalloss;
                        totalloss = 1.f - totalloss;
                        p.slen -= Lmove * prop.y;
                        float ww = currweight - p.w;
                        p.t = tnew;
                        const unsigned int tshift = (unsigned int)gate * gp.framelen;

                        if (gp.outputtype != 2) {                       // :844-851
                            ww = (pd.z == 0.f) ? (currweight * Lmove) : (ww * pd.z);
                        }

                        const bool flushnow = timeup || !isend;

                        if (!GRID) {                                    // :856-1010: run-length merge of deposits into one accumulator
                            const unsigned int newidx = (unsigned int)(p.eid - 1) + tshift;
                            #pragma unroll

                            for (int k = 0; k < 2; k++) {               // k == 1 is the closing flush (one deposit site)
                                const unsigned int idx = (k == 0) ? newidx : (flushnow ? 0xFFFFFFFFu : newidx);

                                if (idx != p.oldidx) {
                                    if (p.oldw > 0.f) {
                                        flush_deposit<false>(gfield, p.oldidx, p.oldw, p, a, hot);
                                    }

                                    p.oldidx = idx;
                                    p.oldw = 0.f;
                                }

                                if (k == 0) {
                                    p.oldw += ww;
                                }
                            }
                        } else {                                        // dual-grid deposit :1022-1206
                            const int seg = ((int)(Lmove * gp.dstep) + 1) << 1;
                            const float seglen = Lmove / seg;
                            const float segdecay = __expf(-prop.x * seglen);
                            const float vs = seglen * gp.dstep;
                            const float dx = p.vx * vs, dy = p.vy * vs, dz = p.vz * vs;
                            float sx = (p.px - gp.nmin[0]) * gp.dstep + dx * 0.5f, sy = (p.py - gp.nmin[1]) * gp.dstep + dy * 0.5f,
                                  sz = (p.pz - gp.nmin[2]) * gp.dstep + dz * 0.5f;
                            const float frac = (totalloss == 0.f) ? 0.f : (1.f - segdecay) / totalloss;
                            float segw = ww;
                            const unsigned int cx = gp.crop0[0], cy = gp.crop0[1];
                            MMCB_UNROLL(MMCB_GRID_UNROLL)

                            for (int k = 0; k < seg; k++) {
                                const int ix = max(__float2int_rd(sx), 0), iy = max(__float2int_rd(sy), 0), iz = max(__float2int_rd(sz), 0);
                                const unsigned int newidx = (unsigned int)iz * cy + (unsigned int)iy * cx + (unsigned int)ix + tshift;

                                if (newidx != p.oldidx) {
                                    if (p.oldw > 0.f) {
                                        flush_deposit<false>(gfield, p.oldidx, p.oldw, p, a, hot);
                                    }

                                    p.oldidx = newidx;
                                    p.oldw = 0.f;
                                }

                                p.oldw += segw * frac;
                                segw *= segdecay;
                                sx += dx;
                                sy += dy;
                                sz += dz;
                            }

                            if (flushnow) {
                                if (p.oldw > 0.f) {
                                    flush_deposit<false>(gfield, p.oldidx, p.oldw, p, a, hot);
                                }

                                p.oldidx = 0xFFFFFFFFu;
                                p.oldw = 0.f;
                            }
                        }
                    }   // found
                }

                if (found) {
                    p.px += Lmove * p.vx;                       // :1222
                    p.py += Lmove * p.vy;
                    p.pz += Lmove * p.vz;
                    p.fixcount = (Lmove > 0.f) ? 0 : (p.fixcount + 0x10000);      // progress guard, see mmcb_photon_kernel

                    if (DET) {                                  // :1943-1945
                        if (type != x.acct) {
                            PPATH_FLUSH();
                            x.acct = type;
                        }

                        x.accL += Lmove;
                    }

                    if (timeup || p.fixcount >= (MMCB_MAX_STALL << 16)) {
                        terminate = true;                       // :1928-1930 / :2007-2009 (photon stays inside: no detection)
                    } else if (!isend) {
                        // ---- cross the face: neighbour hop + boundary physics :1950-1990
                        if (gp.isreflect && (flags & 1u)) {
                            const mmcb_tetrec* rec = a.tet + (p.eid - 1);   // outward normal of the exit face: L1-resident
                            reflectray(p, neweid, __ldg(rec->nx + faceidx), __ldg(rec->ny + faceidx), __ldg(rec->nz + faceidx), prop.w, smed, a, rng);
                        }

                        if (neweid <= 0) {
                            terminate = true;
                            detect = true;
                        } else if (neweid != p.eid) {
                            if ((flags & 0x100u) && !gp.voidtime) {
                                p.t = 0.f;                      // :1970-1978
                            }

                            if ((flags & 0x10u) && !gp.isextdet) {
                                terminate = true;               // :1981-1990 (r.eid = 0)
                                detect = true;
                            } else {
                                p.eid = neweid;
                            }
                        }
                    } else {
                        // ---- end of the scattering path: roulette :2101-2114, then a new direction :2117-2135
                        bool dead = false;

                        if (p.w < gp.roulette_w) {
                            if (rand01(rng) * gp.roulettesize <= 1.f) {
                                p.w *= gp.roulettesize;
                            } else {
                                dead = true;
                            }
                        }

                        if (dead) {
                            terminate = true;
                        } else {
                            float mom;
                            p.slen0 = next_scatter(prop.z, p, rng, mom);
                            p.slen = p.slen0;

                            if (DET) {
                                x.accM += mom;
                                x.accN += 1.f;
                            }
                        }
                    }
                } else {
                    // no exit face found: pull the photon towards the centroid and retry (:1932-1935, :2013-2024)
                    if ((p.fixcount++ & 0xFF) < MMCB_MAX_TRIAL) {
                        const float4 c = a.cent[p.eid - 1];
                        p.px += (c.x - p.px) * FIX_PHOTON;
                        p.py += (c.y - p.py) * FIX_PHOTON;
                        p.pz += (c.z - p.pz) * FIX_PHOTON;
                    } else {
                        terminate = true;                       // dropped without detection
                    }
                }

                if (terminate) {
                    if (DET) {
                        PPATH_FLUSH();
                        x.acct = 0;

                        if (detect) {                           // finddetector :608-621 / wide-field :2072
                            if (gp.isextdet && type == M + 1) {
                                detid = p.eid;
                            } else {
                                for (int i = 0; i < gp.detnum; i++) {
                                    const float4 dp = gdet[i];
                                    const float ddx = dp.x - p.px, ddy = dp.y - p.py, ddz = dp.z - p.pz;

                                    if (ddx * ddx + ddy * ddy + ddz * ddz < dp.w * dp.w) {
                                        detid = i + 1;
                                        break;
                                    }
                                }
                            }
                        }
                    }

                    eesc += p.w;
                    x.mode = RP_NEED;
                } else {
                    classify = true;
                }
            }

            if (DET) {
                const unsigned detmask = __ballot_sync(FULL, detid != 0);

                if (detmask) {                                  // warp-ballot compaction of savedetphoton (:624-682)
                    unsigned int base = 0;

                    if (lane == __ffs(detmask) - 1) {
                        base = atom_add_u32(a.detcount, (unsigned int)__popc(detmask));
                    }

                    base = __shfl_sync(FULL, base, __ffs(detmask) - 1);

                    if (detid) {
                        const unsigned int slot = base + __popc(detmask & ltmask);

                        if (slot < gp.maxdetphoton) {
                            float* out = a.detected + (size_t)slot * (reclen + 1);

                            if (gp.issaveexit) {
                                PPATH(reclen - 7) = p.px;
                                PPATH(reclen - 6) = p.py;
                                PPATH(reclen - 5) = p.pz;
                                PPATH(reclen - 4) = p.vx;
                                PPATH(reclen - 3) = p.vy;
                                PPATH(reclen - 2) = p.vz;
                            }

                            out[0] = (float)(unsigned int)detid;

                            for (int k = 0; k < reclen; k++) {
                                out[1 + k] = PPATH(k);
                            }

                            if (gp.issaveseed) {
                                a.detseed[2 * (size_t)slot] = x.initseed.t0;
                                a.detseed[2 * (size_t)slot + 1] = x.initseed.t1;
                            }
                        }
                    }
                }
            }
        }

        // ------------------------------------------------------------------ classify the walker's next step
        if (classify) {
            const mmcb_tetrec* rec = a.tet + (p.eid - 1);
            float r0[8], r1[8], r2[8];
            ld256(rec, r0);
            ld256((const char*)rec + 32, r1);
            ld256((const char*)rec + 64, r2);
            const int type = __float_as_int(r2[4]);
            const float4 pd = smed[2 * type + 1];
            const float Lmove = (pd.x == 0.f) ? R_MIN_MUS : p.slen * pd.x;
            // the step ends inside iff every face with N.v > 0 lies farther than Lmove: T_j - Lmove S_j > 0 (faces with S_j <= 0 pass
            // as long as the photon is inside); anything doubtful (on a face, outside, degenerate) goes to the general code
            float m = 3.0e38f, smax = -1.f;
            #pragma unroll

            for (int j = 0; j < 4; j++) {
                const float S = p.vx * r0[j] + p.vy * r0[4 + j] + p.vz * r1[j];
                const float Tn = r1[4 + j] - (p.px * r0[j] + p.py * r0[4 + j] + p.pz * r1[j]);
                m = fminf(m, Tn - Lmove * S);
                smax = fmaxf(smax, S);
            }

            const int gate = (int)((p.t + Lmove * pd.y - gp.tstart) * gp.Rtstep);
            bool fast = (m > 0.f) && (smax > 0.f) && (Lmove > 0.f) && (gate <= gp.maxgate - 1) && (gate >= 0);
            const unsigned int tshift = (unsigned int)gate * gp.framelen;
            unsigned int newidx;

            if (!GRID) {
                newidx = (unsigned int)(p.eid - 1) + tshift;
            } else {
                // two segments (Lmove shorter than a voxel edge), both midpoints in one voxel: the arithmetic of the general loop
                const float vs = (Lmove * 0.5f) * gp.dstep;
                const float dx = p.vx * vs, dy = p.vy * vs, dz = p.vz * vs;
                const float sx = (p.px - gp.nmin[0]) * gp.dstep + dx * 0.5f, sy = (p.py - gp.nmin[1]) * gp.dstep + dy * 0.5f,
                            sz = (p.pz - gp.nmin[2]) * gp.dstep + dz * 0.5f;
                const int ix = max(__float2int_rd(sx), 0), iy = max(__float2int_rd(sy), 0), iz = max(__float2int_rd(sz), 0);
                const int jx = max(__float2int_rd(sx + dx), 0), jy = max(__float2int_rd(sy + dy), 0), jz = max(__float2int_rd(sz + dz), 0);
                fast = fast && (Lmove * gp.dstep < 1.f) && (ix == jx) && (iy == jy) && (iz == jz);
                newidx = (unsigned int)iz * gp.crop0[1] + (unsigned int)iy * gp.crop0[0] + (unsigned int)ix + tshift;
            }

            x.nidx = newidx;
            x.type = type;
            x.mode = fast ? RP_F : RP_G;
        }
    }

#undef PPATH_FLUSH
#undef PPATH
    // stream states go back in the seed-word packing (register walker and stashed walker of every lane)
    __syncwarp();
    // (which walker ends in which lane or slot is a permutation of the warp's 64 streams; every state is written exactly once)
    *(uint4*)(a.seeds + 4 * (gw * 64 + lane)) = make_uint4((unsigned int)(rng.t0 >> 32), (unsigned int)rng.t0, (unsigned int)(rng.t1 >> 32), (unsigned int)rng.t1);
    *(uint4*)(a.seeds + 4 * (gw * 64 + 32 + lane)) = stash[3 * MMCB_RP_SLOTS + lane];

    if (hoton) {            // flush the CTA-private sums of the hot lines
        __syncthreads();

        for (int i = threadIdx.x; i < MMCB_HOT_SLOTS * MMCB_HOT_GROUP; i += blockDim.x) {
            const unsigned int g = hkeys[i >> MMCB_HOT_GROUP_LOG2];
            const float v = hvals[i];
            const unsigned int idx = (g << MMCB_HOT_GROUP_LOG2) + (i & (MMCB_HOT_GROUP - 1));

            if (g != MMCB_HOT_EMPTY && v != 0.f && idx < gp.fieldlen) {
                red_add(field + idx, v);
            }
        }
    }

    double dt = etot, de = eesc, dr = (double)nraytet;
    #pragma unroll

    for (int o = 16; o > 0; o >>= 1) {
        dt += __shfl_xor_sync(FULL, dt, o);
        de += __shfl_xor_sync(FULL, de, o);
        dr += __shfl_xor_sync(FULL, dr, o);
    }

    if (lane == 0) {
        red_add_d(a.energy, dt);
        red_add_d(a.energy + MMCB_MAX_SRCNUM, de);
        red_add_d(a.raytet, dr);
    }
}
