// P2G, production kernel: lane <-> cell COLUMN, marching along z -- no atomics, fixed summation order
// (bit-reproducible), every particle field read from global memory as coalesced 8-byte async copies that are
// issued one window ahead of the arithmetic.
// Reference: simulation::_transfer_to_grid_{pic,flip,apic}, src/simulation.cpp:293-412, 428-445, 72-78.
//
// A block owns the +faces of PM_BX x (WARPS - 2) cell columns over a chunk of z planes.  Warp <-> one row y of 32 cells
// (the WARPS - 2 owned rows plus one halo row either side), lane <-> cell x (30 owned columns plus one halo column either
// side).  All warps march through the cell layers of the chunk together.  For the layer it is in, a lane runs over
// the particles of ITS cell and accumulates, in registers, their contributions to the 2 x 3 x 3 faces the cell can
// reach (per velocity component) -- the per-cell arithmetic of p2g_accum.cuh.  Where the sums go next:
//   * along z the accumulators ROTATE: the partial sum a column holds for face plane k collects the contributions
//     of its cells in layers k-1, k, k+1 without leaving the registers, and is flushed once, when the march has
//     passed layer k+1 (3x fewer flushed values than one flush per cell);
//   * along x the three (two) columns that reach a face are combined with warp shuffles;
//   * along y each warp writes its three (two) row partials into its own shared-memory slot and ANNOUNCES them on an
//     mbarrier; one layer later it waits for the announcements of all warps, and the warp that owns row y adds
//     slot[y-1], slot[y], slot[y+1] in that fixed order, normalises, classifies, zeroes boundary faces, takes the
//     FLIP snapshot, adds gravity and writes the finished face row.  Two slot sets alternate, so a warp whose rows
//     hold fewer particles runs up to one layer ahead of the others instead of idling at a block barrier per layer
//     (the __syncthreads() form: 15.3 ms at 256^3, this form 14.75, r3k).
// No colouring, no read-modify-write on shared accumulators.  Halo recomputation: 32/30 in x, WARPS/(WARPS - 2) in y,
// (chunk + 2)/chunk in z (1.33x .. 1.4x).
//
// Staging: per warp two buffers of [field][cell * PB_CSTRIDE + slot]; window n+1 (4 particle slots per cell) is
// copied with cp.async while window n is being accumulated, and the permutation indices (lean sort: velocity / c
// rows are still in pre-sort order) of window n+2 are loaded into registers at the same time, so no global-memory
// latency sits between two windows of arithmetic.
//
// This file is compiled WITH fused multiply-add (P2G is tolerance-checked: rel-L2 <= 1e-12 against the oracle).
#include "lfk_internal.cuh"
#include "p2g_accum.cuh"

#define PM_BX 30
// WARPS (template parameter): warps per block = rows per block (two of them halo rows).  Warps are placed on the 4 SM
// sub-partitions (16 K registers each), so ceil(WARPS / 4) * 32 * registers must fit: 8 warps at 200 registers, or 10 at 168.
#ifndef PM_REGS8
#define PM_REGS8 216
#endif
#define PM_REGS(WARPS) ((WARPS) <= 8 ? PM_REGS8 : 168)
#define PM_STAGE (PB_FIELDS * PB_FSTRIDE)       // doubles per staging buffer
#define PM_SLOT (3 * 2 * 32)                    // doubles per warp slot: [b][w | wv][lane]
#define PM_MAX_CHUNK 128                        // z planes per block at most (size of the z-coordinate table)
#define PM_FULL 0xffffffffu

struct PMWin { // one staging window: WIN particle slots of every cell of the warp's row in layer lz
	int lz, win, maxcnt, valid;
	uint32_t pb, pe; // this lane's cell: particle range
};

__device__ __forceinline__ void cp_async_commit() {
	asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_1() {
	asm volatile("cp.async.wait_group 1;" ::: "memory");
}
// split block barrier (mbarrier): a warp announces that its row partials of a layer are in shared memory and only waits
// when it needs its neighbours' -- one layer later -- so warps whose rows hold fewer particles run ahead instead of
// idling at a __syncthreads() per layer
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
	asm volatile("{\n"
		".reg .pred p;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra LAB_DONE;\n"
		"bra LAB_WAIT;\n"
		"LAB_DONE:\n"
		"}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
struct PMSync { // per-thread view of the block's split barrier
	unsigned long long *bar;
	unsigned phase; // parity of the next phase to wait for
	int pend;       // local layer index of the face plane whose row partials are announced but not yet combined (-1: none)
};
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { v = max(v, __shfl_xor_sync(PM_FULL, v, o)); }
	return v;
}

// per-warp marching state that is not component specific
struct PMRow {
	const uint32_t *__restrict__ begin;
	long long row0, rstride; // raw index of (x, y, lz = 0), and nx * ny
	int ok;                  // this lane's cell is inside the grid in x and y
	int lzB;                 // last layer of the march
	uint32_t pref_pb, pref_pe;
	// (no select on the loaded values -- a consumer right behind the load would stall the warp for the whole memory
	// latency: a lane outside the grid reads the same entry twice and gets an empty range)
	__device__ __forceinline__ void load(int lz, uint32_t &pb, uint32_t &pe) const {
		const long long c = ok ? row0 + rstride * lz : 0;
		pb = begin[c];
		pe = begin[c + ok];
	}
	// next window of the stream; every layer has at least one (possibly empty) window
	__device__ __forceinline__ void advance(PMWin &D) {
		if (!D.valid) { return; }
		if (D.win + PB_WIN < D.maxcnt) {
			D.win += PB_WIN;
			return;
		}
		if (D.lz >= lzB) {
			D.valid = 0;
			return;
		}
		D.lz += 1;
		D.win = 0;
		D.pb = pref_pb; // loaded one layer ago
		D.pe = pref_pe;
		D.maxcnt = warp_max_i((int)(D.pe - D.pb));
		pref_pb = 0;
		pref_pe = 0;
		if (D.lz + 1 <= lzB) { load(D.lz + 1, pref_pb, pref_pe); }
	}
};

// source indices of the velocity / c rows of the window's particles (lean sort), one window ahead of their use
__device__ __forceinline__ void pm_perm_load(const PMWin &D, const uint32_t *__restrict__ perm, int lane,
	uint32_t *R) {
#pragma unroll
	for (int k = 0; k < PB_WIN; ++k) {
		const int t = k * 32 + lane, sc = t / PB_WIN, ss = t % PB_WIN;
		const uint32_t qb = __shfl_sync(PM_FULL, D.pb, sc), qe = __shfl_sync(PM_FULL, D.pe, sc);
		const uint32_t q = qb + (uint32_t)D.win + (uint32_t)ss;
		// slots past the end of their cell read entry 0 (their index is never used): the loaded value goes straight
		// into R[k], nothing waits for it before the next window's pm_issue (r3l: a select behind the load stalled the
		// warp for the memory latency once per window, 13 % of the kernel's stall samples)
		const uint32_t qs = (D.valid && q < qe) ? q : 0u;
		R[k] = perm != nullptr ? perm[qs] : q;
	}
}

// lanes <-> consecutive particles on the global side (coalesced), [cell][slot] on the shared side
template <int COMP, bool APIC> __device__ __forceinline__ void pm_issue(const PMWin &D, const uint32_t *R,
	double *__restrict__ st, const double *const *fields, int lane) {
	if (D.valid) {
#pragma unroll
		for (int k = 0; k < PB_WIN; ++k) {
			const int t = k * 32 + lane, sc = t / PB_WIN, ss = t % PB_WIN;
			const uint32_t qb = __shfl_sync(PM_FULL, D.pb, sc), qe = __shfl_sync(PM_FULL, D.pe, sc);
			const uint32_t q = qb + (uint32_t)D.win + (uint32_t)ss;
			if (q < qe) {
				double *dst = st + sc * PB_CSTRIDE + ss;
#pragma unroll
				for (int f = 0; f < 3; ++f) { cp_async8(dst + f * PB_FSTRIDE, fields[f] + q); }
				const uint32_t qv = R[k];
				cp_async8(dst + 3 * PB_FSTRIDE, fields[PF_VX + COMP] + qv);
				if (APIC) {
#pragma unroll
					for (int f = 0; f < 3; ++f) { cp_async8(dst + (4 + f) * PB_FSTRIDE, fields[PF_C0 + 3 * COMP + f] + qv); }
				}
			}
		}
	}
	cp_async_commit();
}

struct PMOut { // where one component goes
	double *__restrict__ out, *__restrict__ old;
	uint8_t *__restrict__ typ;
};

// The row partials of face plane `op` are complete in slot set `set` on every warp: the warp that owns row y adds
// slot[y-1], slot[y], slot[y+1] in that fixed order, normalises, classifies, zeroes boundary faces, takes the FLIP
// snapshot, adds gravity and writes the finished face row.
template <int COMP, int METHOD, int WARPS> __device__ __forceinline__ void pm_store_plane(const GridDesc &G, const PBParams &Q,
	int op, const double *__restrict__ set, int warp, int lane, int x, int y, bool owner, uint8_t type_now,
	uint32_t count_now, const PMOut &O) {
	constexpr int NB = COMP == 1 ? 2 : 3;
	constexpr bool APIC = METHOD == LFK_METHOD_APIC;
	if (owner) {
		const double *mine = set + warp * PM_SLOT;
		const double *below = set + (warp - 1) * PM_SLOT;
		const double *above = set + (warp + 1) * PM_SLOT;
		double sw, sv;
		if (NB == 3) { // row y-1 (b = 2), row y (b = 1), row y+1 (b = 0)
			sw = below[(2 * 2 + 0) * 32 + lane];
			sv = below[(2 * 2 + 1) * 32 + lane];
			sw += mine[(1 * 2 + 0) * 32 + lane];
			sv += mine[(1 * 2 + 1) * 32 + lane];
			sw += above[(0 * 2 + 0) * 32 + lane];
			sv += above[(0 * 2 + 1) * 32 + lane];
		} else { // staggered in y: row y (b = 1), row y+1 (b = 0)
			sw = mine[(1 * 2 + 0) * 32 + lane];
			sv = mine[(1 * 2 + 1) * 32 + lane];
			sw += above[(0 * 2 + 0) * 32 + lane];
			sv += above[(0 * 2 + 1) * 32 + lane];
		}
		const int z = op - 1 + G.z0;
		const long long me = x + (long long)G.nx * (y + (long long)G.ny * op);
		double r = sw > 1e-6 ? sv / sw : 0.0; // src/simulation.cpp:380-386
		const bool edge = COMP == 0 ? x == G.nx - 1 : (COMP == 1 ? y == G.ny - 1 : z == G.nz - 1);
		if (METHOD == LFK_METHOD_FLIP) { O.old[me] = edge ? 0.0 : r; } // :340-344
		if (APIC && edge) { r = 0.0; }                                  // :397
		if (Q.add_gravity) { r += Q.gdt[COMP]; }                        // :72-78
		O.out[me] = r;
		if (COMP == 0) { // classification, once per cell (:388-393)
			if (type_now != LFK_CELL_SOLID) { O.typ[me] = count_now > 0 ? LFK_CELL_FLUID : LFK_CELL_AIR; }
		}
	}
}

// the announced plane, if any: wait until every warp has announced it, then combine and write it
template <int COMP, int METHOD, int WARPS> __device__ __forceinline__ void pm_complete_pending(const GridDesc &G, const PBParams &Q,
	const double *__restrict__ slots, int par, PMSync &S, int warp, int lane, int x, int y, int nfx,
	const uint32_t *__restrict__ begin, const PMOut &O) {
	if (S.pend < 0) { return; } // block-uniform
	const int fx = lane - 1; // owned face columns: lanes 1 .. 30
	const bool owner = warp >= 1 && warp <= WARPS - 2 && y < G.ny && fx >= 0 && fx < nfx;
	// the cell's type and particle count (classification) are requested before the wait, not behind it
	uint8_t type_now = LFK_CELL_SOLID;
	uint32_t b0 = 0, b1 = 0;
	if (COMP == 0 && owner) {
		const long long me = x + (long long)G.nx * (y + (long long)G.ny * S.pend);
		type_now = O.typ[me];
		b0 = begin[me];
		b1 = begin[me + 1];
	}
	mbar_wait(S.bar, S.phase);
	S.phase ^= 1u;
	pm_store_plane<COMP, METHOD, WARPS>(G, Q, S.pend, slots + (par ^ 1) * WARPS * PM_SLOT, warp, lane, x, y, owner, type_now,
		b1 - b0, O);
	S.pend = -1;
}

// Layer L of the march is complete: face plane L - 1 has everything this column contributes.  Combine along x
// (shuffles), announce the row partials (slots; combined along y one layer later), rotate the accumulators.
template <int COMP, int METHOD, int WARPS> __device__ __forceinline__ void pm_finish_layer(const GridDesc &G, const PBParams &Q,
	int L, int P0, int P1, double *__restrict__ slots, int &par, PMSync &S, int warp, int lane, int x, int y, int nfx,
	const uint32_t *__restrict__ begin, const PMOut &O, double *accw, double *accv) {
	constexpr int NA = COMP == 0 ? 2 : 3, NB = COMP == 1 ? 2 : 3, NC = COMP == 2 ? 2 : 3;
	const int op = L - 1; // local layer index of the finished face plane
	if (op >= P0 && op < P1) { // block-uniform
		// the plane announced one layer ago used the other slot set; once every warp has announced it, nobody reads
		// the set written below any more (its last readers combined it before they announced that plane)
		pm_complete_pending<COMP, METHOD, WARPS>(G, Q, slots, par, S, warp, lane, x, y, nfx, begin, O);
		double *mine = slots + (par * WARPS + warp) * PM_SLOT;
#pragma unroll
		for (int b = 0; b < NB; ++b) {
			double rw, rv;
			if (NA == 3) { // face of cell X: column X-1 (a = 2), column X (a = 1), column X+1 (a = 0)
				rw = __shfl_up_sync(PM_FULL, accw[b * NA + 2], 1);
				rv = __shfl_up_sync(PM_FULL, accv[b * NA + 2], 1);
				rw += accw[b * NA + 1];
				rv += accv[b * NA + 1];
				rw += __shfl_down_sync(PM_FULL, accw[b * NA + 0], 1);
				rv += __shfl_down_sync(PM_FULL, accv[b * NA + 0], 1);
			} else { // staggered in x: column X (a = 1), column X+1 (a = 0)
				rw = accw[b * NA + 1];
				rv = accv[b * NA + 1];
				rw += __shfl_down_sync(PM_FULL, accw[b * NA + 0], 1);
				rv += __shfl_down_sync(PM_FULL, accv[b * NA + 0], 1);
			}
			mine[(b * 2 + 0) * 32 + lane] = rw;
			mine[(b * 2 + 1) * 32 + lane] = rv;
		}
		mbar_arrive(S.bar); // release: this thread's partials are visible to whoever completes the phase
		S.pend = op;
		par ^= 1;
	}
	// rotate: plane c+1 of this layer is plane c of the next one
#pragma unroll
	for (int t = 0; t < (NC - 1) * NB * NA; ++t) {
		accw[t] = accw[t + NB * NA];
		accv[t] = accv[t + NB * NA];
	}
#pragma unroll
	for (int t = (NC - 1) * NB * NA; t < NC * NB * NA; ++t) {
		accw[t] = 0.0;
		accv[t] = 0.0;
	}
}

template <int COMP, int METHOD, int WARPS> __device__ __forceinline__ void pm_march(const GridDesc &G, const PBParams &Q,
	double *__restrict__ st, double *__restrict__ slots, const double *__restrict__ ztab, int &par, PMSync &S,
	const double *const *fields, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ begin,
	const double *__restrict__ cxs, const double *__restrict__ cys, int warp, int lane, int x0, int y0, int P0, int P1,
	int nfx, const PMOut &O) {
	constexpr bool APIC = METHOD == LFK_METHOD_APIC;
	const int x = x0 - 1 + lane, y = y0 - 1 + warp;
	const bool xin = x >= 0 && x < G.nx, yin = y >= 0 && y < G.ny;
	const int lzA = P0 - 1, lzB = P1; // P0 >= 1, P1 <= nzl + 1 = nlz - 1
	// cell-centre coordinates of cell - 1, cell, cell + 1 per axis (tables built by repeated addition like the
	// reference; entries outside the grid are extrapolated by +-h and only ever carry targets outside the grid)
	double cc[9];
	{
		const int xm = x - 1, xp = x + 1;
		const double xc = xin ? cxs[x] : (x < 0 ? cxs[0] - G.h : cxs[G.nx - 1] + G.h);
		cc[0] = (xm >= 0 && xm < G.nx) ? cxs[xm] : xc - G.h;
		cc[1] = xc;
		cc[2] = (xp >= 0 && xp < G.nx) ? cxs[xp] : xc + G.h;
		const int ym = y - 1, yp = y + 1;
		const double yc = yin ? cys[y] : (y < 0 ? cys[0] - G.h : cys[G.ny - 1] + G.h);
		cc[3] = (ym >= 0 && ym < G.ny) ? cys[ym] : yc - G.h;
		cc[4] = yc;
		cc[5] = (yp >= 0 && yp < G.ny) ? cys[yp] : yc + G.h;
		cc[6] = cc[7] = cc[8] = 0.0;
	}
	double accw[18], accv[18];
#pragma unroll
	for (int t = 0; t < 18; ++t) {
		accw[t] = 0.0;
		accv[t] = 0.0;
	}
	PMRow R;
	R.begin = begin;
	R.rstride = G.sxy;
	R.row0 = (long long)G.nx * (yin ? y : 0) + (xin ? x : 0);
	R.ok = xin && yin;
	R.lzB = lzB;
	R.pref_pb = R.pref_pe = 0;

	// ---- prologue: window 0 in flight, permutation indices of window 1 loaded ----
	PMWin D0, D1, D2;
	uint32_t R1[PB_WIN], R2[PB_WIN];
	D2.lz = lzA;
	D2.win = 0;
	D2.valid = 1;
	R.load(lzA, D2.pb, D2.pe);
	D2.maxcnt = warp_max_i((int)(D2.pe - D2.pb));
	if (lzA + 1 <= lzB) { R.load(lzA + 1, R.pref_pb, R.pref_pe); }
	pm_perm_load(D2, perm, lane, R1);
	D1 = D2;
	R.advance(D2);
	pm_issue<COMP, APIC>(D1, R1, st, fields, lane);
	pm_perm_load(D2, perm, lane, R1);
	D0 = D1;
	D1 = D2;
	R.advance(D2);
	int cur = 0;
	while (D0.valid) { // warp-uniform
		pm_perm_load(D2, perm, lane, R2);
		pm_issue<COMP, APIC>(D1, R1, st + (cur ^ 1) * PM_STAGE, fields, lane);
		cp_async_wait_1(); // everything but the window just issued has landed: window D0 is in st[cur]
		__syncwarp();
		if (D0.win == 0) { // first window of a layer
			if (D0.lz > lzA) {
				pm_finish_layer<COMP, METHOD, WARPS>(G, Q, D0.lz - 1, P0, P1, slots, par, S, warp, lane, x, y, nfx, begin, O, accw, accv);
			}
			const int zi = D0.lz - lzA; // ztab[i]: centre of the layer (lzA - 1 + i)
			cc[6] = ztab[zi];
			cc[7] = ztab[zi + 1];
			cc[8] = ztab[zi + 2];
		}
		const int cnt = (int)(D0.pe - D0.pb);
		const int nslots = max(0, min(PB_WIN, cnt - D0.win));
		accumulate_cell<COMP, APIC>(st + cur * PM_STAGE, lane, nslots, cc, Q.half, Q.inv_h, accw, accv);
		__syncwarp(); // st[cur] is overwritten by the window after next
		D0 = D1;
		D1 = D2;
#pragma unroll
		for (int k = 0; k < PB_WIN; ++k) { R1[k] = R2[k]; }
		R.advance(D2);
		cur ^= 1;
	}
	pm_finish_layer<COMP, METHOD, WARPS>(G, Q, lzB, P0, P1, slots, par, S, warp, lane, x, y, nfx, begin, O, accw, accv);
	pm_complete_pending<COMP, METHOD, WARPS>(G, Q, slots, par, S, warp, lane, x, y, nfx, begin, O);
	cp_async_wait_all();
}

template <int METHOD, int WARPS> __global__ void __maxnreg__(PM_REGS(WARPS)) k_p2g_march(GridDesc G, PBParams Q,
	ParticleSoA P, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ begin,
	const double *__restrict__ cxs, const double *__restrict__ cys, const double *__restrict__ czs,
	double *__restrict__ u, double *__restrict__ v, double *__restrict__ w, double *__restrict__ uo,
	double *__restrict__ vo, double *__restrict__ wo, uint8_t *__restrict__ typ, int chunk) {
	extern __shared__ double smem[];
	double *stage_all = smem;                                   // [WARPS][2][PM_STAGE]
	double *slots = smem + WARPS * 2 * PM_STAGE;                // [2][WARPS][PM_SLOT]
	double *ztab = slots + 2 * WARPS * PM_SLOT;                 // [PM_MAX_CHUNK + 4]
	__shared__ unsigned long long layer_bar;
	if (threadIdx.x == 0) { mbar_init(&layer_bar, WARPS * 32); }
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	double *st = stage_all + warp * (2 * PM_STAGE);
	const int x0 = blockIdx.x * PM_BX, y0 = blockIdx.y * (WARPS - 2);
	const int P0 = 1 + blockIdx.z * chunk, P1 = min(P0 + chunk, G.nzl + 1);
	const int nfx = min(PM_BX, G.nx - x0);
	// z centres of the layers lzA - 1 .. lzB + 1 (global z = local layer - 1 + z0), extrapolated outside the grid
	for (int i = threadIdx.x; i < P1 - P0 + 4; i += WARPS * 32) {
		const int z = (P0 - 2 + i) - 1 + G.z0;
		ztab[i] = z < 0 ? czs[0] + (double)z * G.h : (z >= G.nz ? czs[G.nz - 1] + (double)(z - G.nz + 1) * G.h : czs[z]);
	}
	__syncthreads();
	const double *fields[15];
#pragma unroll
	for (int f = 0; f < 15; ++f) { fields[f] = P.f[f]; }
	int par = 0;
	PMSync S{ &layer_bar, 0u, -1 };
	const PMOut O0{ u, uo, typ }, O1{ v, vo, typ }, O2{ w, wo, typ };
	pm_march<0, METHOD, WARPS>(G, Q, st, slots, ztab, par, S, fields, perm, begin, cxs, cys, warp, lane, x0, y0, P0, P1, nfx, O0);
	pm_march<1, METHOD, WARPS>(G, Q, st, slots, ztab, par, S, fields, perm, begin, cxs, cys, warp, lane, x0, y0, P0, P1, nfx, O1);
	pm_march<2, METHOD, WARPS>(G, Q, st, slots, ztab, par, S, fields, perm, begin, cxs, cys, warp, lane, x0, y0, P0, P1, nfx, O2);
}

int lfkp_materialise_vc(lfk_ctx *c);

template <int WARPS> static int p2g_march_launch(lfk_ctx *c, const PBParams &Q, const uint32_t *perm) {
	const GridDesc &G = c->g;
	const unsigned nbx = (unsigned)((G.nx + PM_BX - 1) / PM_BX), nby = (unsigned)((G.ny + (WARPS - 2) - 1) / (WARPS - 2));
	// z chunks: enough blocks for ~6 waves of one block per SM, chunks of 8 .. PM_MAX_CHUNK planes
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
	int nzc = (int)((6u * (unsigned)sms + nbx * nby - 1) / (nbx * nby));
	int chunk = c->tune.p2g_chunk > 0 ? c->tune.p2g_chunk : (G.nzl + nzc - 1) / nzc;
	chunk = chunk < 8 ? 8 : chunk;
	chunk = chunk > PM_MAX_CHUNK ? PM_MAX_CHUNK : chunk;
	chunk = chunk > G.nzl ? G.nzl : chunk;
	nzc = (G.nzl + chunk - 1) / chunk;
	dim3 grid(nbx, nby, (unsigned)nzc);
	const size_t smem = (size_t)(WARPS * 2 * PM_STAGE + 2 * WARPS * PM_SLOT + PM_MAX_CHUNK + 4) * sizeof(double);
	static bool attr_set[LFK_MAX_DEVICES] = {}; // function attributes are per device
	if (!attr_set[c->device % LFK_MAX_DEVICES]) {
		LFK_CUDA(c, cudaFuncSetAttribute((k_p2g_march<LFK_METHOD_PIC, WARPS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LFK_CUDA(c, cudaFuncSetAttribute((k_p2g_march<LFK_METHOD_FLIP, WARPS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LFK_CUDA(c, cudaFuncSetAttribute((k_p2g_march<LFK_METHOD_APIC, WARPS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set[c->device % LFK_MAX_DEVICES] = true;
	}
	switch (c->prm.method) {
	case LFK_METHOD_PIC:
		LFK_LAUNCH(c, (k_p2g_march<LFK_METHOD_PIC, WARPS>), grid, WARPS * 32, smem, G, Q, c->P, perm, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ, chunk);
		break;
	case LFK_METHOD_FLIP:
		LFK_LAUNCH(c, (k_p2g_march<LFK_METHOD_FLIP, WARPS>), grid, WARPS * 32, smem, G, Q, c->P, perm, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ, chunk);
		break;
	default:
		LFK_LAUNCH(c, (k_p2g_march<LFK_METHOD_APIC, WARPS>), grid, WARPS * 32, smem, G, Q, c->P, perm, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ, chunk);
		break;
	}
	return 0;
}

int lfkg_p2g_march(lfk_ctx *c, double gravity_dt, bool add_gravity) {
	const GridDesc &G = c->g;
	PBParams Q;
	Q.half = 0.5 * G.h;
	Q.inv_h = 1.0 / G.h;
	Q.add_gravity = add_gravity ? 1 : 0;
	for (int d = 0; d < 3; ++d) {
		Q.gdt[d] = c->prm.gravity[d] * gravity_dt;
	}
	// one permutation serves the velocity and the c rows: make them agree (they differ only after lfkp_permute_c)
	if (c->prm.method == LFK_METHOD_APIC && c->v_deferred != c->c_deferred) { LFK_TRY(lfkp_materialise_vc(c)); }
	const uint32_t *perm = c->v_deferred ? c->perm : nullptr;
	// (10 rows per block at 168 registers spill the accumulators: 25.8 ms against 15.2 ms at 256^3, r2c sweep; 8 rows:
	// 200 registers 14.75 ms, 224 (no spills) 14.05, 255 14.05, r3k; after the r3l load fixes 208: 13.55, 216: 12.73,
	// 224: 13.07, 232: 13.06, 255: 13.10, r3o -- register allocation luck decides at this level)
	return p2g_march_launch<8>(c, Q, perm);
}
