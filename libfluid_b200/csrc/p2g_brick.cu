// P2G, production kernel: cell-owner scatter on bricks -- no atomics, fixed summation order (bit-reproducible),
// each particle field read from global memory as coalesced 8-byte async copies.
// Reference: simulation::_transfer_to_grid_{pic,flip,apic}, src/simulation.cpp:293-412, 428-445, 72-78.
//
// A block owns the +faces of a brick of BX x BY x BZ cells and visits every cell that can touch one of them (the
// brick plus a one-cell halo: 32 x 10 x 10 cells).  A warp takes one row of 32 cells, lane <-> cell:
//   1. the row's particles are staged into the warp's shared-memory slab with cp.async -- lanes <-> consecutive
//      particles on the global side (coalesced), transposed to [cell][slot] on the shared side;
//   2. each lane runs over the particles of ITS cell and accumulates, in registers, their contribution to the
//      2 x 3 x 3 faces (per velocity component) the cell can reach;
//   3. the lane adds its 18 partial sums into the block's accumulator tile one neighbour offset at a time.  For a
//      fixed offset all lanes of a row hit distinct faces, and rows are scheduled in 9 colours (y mod 3, z mod 3)
//      so that concurrently active warps never touch the same face: plain read-modify-write, no atomics, and the
//      order of additions does not depend on timing.
// After the last colour the tile holds sum(w) and sum(w v) of every face of the brick, complete: the block
// normalises, classifies, zeroes boundary faces, takes the FLIP snapshot, adds gravity and writes the faces.
// The halo is recomputed by the neighbouring bricks (1.67x redundant flops) -- that buys determinism and removes
// the global accumulator arrays, their zeroing and a separate normalisation pass.
//
// This file is compiled WITH fused multiply-add (the summation order already differs from the reference's, P2G is
// tolerance-checked: rel-L2 <= 1e-12 against the oracle).
#include "lfk_internal.cuh"
#include "p2g_accum.cuh"

// Brick 30 x 10 x 7 faces => 12 x 9 rows of 32 cells; every colour (y mod 3, z mod 3) has exactly 4 x 3 = 12 rows =
// two balanced rounds for 6 warps.  ~90 KB of shared memory per block => two blocks (12 warps) per SM.
#define PB_BX 30
#define PB_BY 10
#define PB_BZ 7
#define PB_RY (PB_BY + 2)
#define PB_RZ (PB_BZ + 2)
#define PB_WARPS 6
#define PB_THREADS (PB_WARPS * 32)
#define PB_ACC_COMP (2 * PB_BZ * PB_BY * 32) // one velocity component: sum(w) and sum(w v) per face

// add the lane's 18 partial sums into the block tile, one neighbour offset at a time
template <int COMP> __device__ __forceinline__ void flush_cell(double *__restrict__ acc, int fx0, int fy0, int fz0,
	int nfx, const double *accw, const double *accv) {
	constexpr int NA = COMP == 0 ? 2 : 3, NB = COMP == 1 ? 2 : 3, NC = COMP == 2 ? 2 : 3;
	// fx0 / fy0 / fz0: tile coordinates of the face owned by (cell - 1) along each axis
#pragma unroll
	for (int c = 0; c < NC; ++c) {
#pragma unroll
		for (int b = 0; b < NB; ++b) {
#pragma unroll
			for (int a = 0; a < NA; ++a) {
				const int fx = fx0 + a, fy = fy0 + b, fz = fz0 + c;
				const int t = (c * NB + b) * NA + a;
				if (fx >= 0 && fx < nfx && fy >= 0 && fy < PB_BY && fz >= 0 && fz < PB_BZ) {
					double *p = acc + (fz * PB_BY + fy) * 32 + fx;
					p[0] += accw[t];
					p[PB_BZ * PB_BY * 32] += accv[t];
				}
				__syncwarp(); // lanes c (offset a) and c + 1 (offset a - 1) share a face: keep the steps ordered
			}
		}
	}
}

// one velocity component of one row: stage, accumulate over all slot windows, flush once
template <int COMP, bool APIC> __device__ __forceinline__ void row_component(double *__restrict__ st,
	double *__restrict__ acc, const double *const *fields, const uint32_t *__restrict__ permv,
	const uint32_t *__restrict__ permc, int lane, uint32_t pb, uint32_t pe, int cnt, int maxcnt,
	const double *cc, const PBParams &Q, int ry, int rz, int nfx) {
	double accw[18], accv[18];
#pragma unroll
	for (int t = 0; t < 18; ++t) {
		accw[t] = 0.0;
		accv[t] = 0.0;
	}
	for (int win = 0; win < maxcnt; win += PB_WIN) {
		const int nslots = max(0, min(PB_WIN, cnt - win));
		// lanes <-> consecutive particles on the global side (coalesced), [cell][slot] on the shared side
#pragma unroll
		for (int k = 0; k < PB_WIN; ++k) {
			const int t = k * 32 + lane, sc = t / PB_WIN, ss = t % PB_WIN;
			const uint32_t qb = __shfl_sync(0xffffffffu, pb, sc), qe = __shfl_sync(0xffffffffu, pe, sc);
			const uint32_t q = qb + win + ss;
			if (q < qe) {
				double *dst = st + sc * PB_CSTRIDE + ss;
#pragma unroll
				for (int f = 0; f < 3; ++f) { cp_async8(dst + f * PB_FSTRIDE, fields[f] + q); }
				// lean sort: velocity / c rows may still be in the pre-sort order (index perm[q])
				const uint32_t qv = permv ? permv[q] : q;
				cp_async8(dst + 3 * PB_FSTRIDE, fields[PF_VX + COMP] + qv);
				if (APIC) {
					const uint32_t qc = permc ? permc[q] : q;
#pragma unroll
					for (int f = 0; f < 3; ++f) { cp_async8(dst + (4 + f) * PB_FSTRIDE, fields[PF_C0 + 3 * COMP + f] + qc); }
				}
			}
		}
		cp_async_wait_all();
		__syncwarp();
		accumulate_cell<COMP, APIC>(st, lane, nslots, cc, Q.half, Q.inv_h, accw, accv);
		__syncwarp(); // the slab is overwritten by the next window
	}
	// tile coordinates of the faces owned by (cell - 1): x: lane - 2, y: ry - 2, z: rz - 2
	flush_cell<COMP>(acc, lane - 2, ry - 2, rz - 2, nfx, accw, accv);
}

template <int COMP, bool APIC> __device__ __forceinline__ void brick_component(const GridDesc &G, const PBParams &Q,
	double *__restrict__ st, double *__restrict__ acc, const double *const *fields,
	const uint32_t *__restrict__ permv, const uint32_t *__restrict__ permc, const uint32_t *__restrict__ begin,
	const double *__restrict__ cxs, const double *__restrict__ cys, const double *__restrict__ czs, int warp, int lane,
	int x0, int y0, int lz0, int nfx) {
	const int x = x0 - 1 + lane; // this lane's cell column
	const bool xin = x >= 0 && x < G.nx;
	// cell-centre coordinates of cell - 1, cell, cell + 1 per axis (table built by repeated addition like the
	// reference; entries outside the grid are extrapolated by +-h and only ever carry targets outside the grid)
	double cc[9];
	{
		const int xm = x - 1, xp = x + 1;
		const double xc = xin ? cxs[x] : (x < 0 ? cxs[0] - G.h : cxs[G.nx - 1] + G.h);
		cc[0] = (xm >= 0 && xm < G.nx) ? cxs[xm] : xc - G.h;
		cc[1] = xc;
		cc[2] = (xp >= 0 && xp < G.nx) ? cxs[xp] : xc + G.h;
	}
	for (int colour = 0; colour < 9; ++colour) {
		const int cy = colour % 3, cz = colour / 3;
		const int nry = (PB_RY - cy + 2) / 3, nrz = (PB_RZ - cz + 2) / 3; // rows cy, cy + 3, ... / cz, cz + 3, ...
		for (int job = warp; job < nry * nrz; job += PB_WARPS) {
			const int ry = cy + 3 * (job % nry), rz = cz + 3 * (job / nry);
			const int y = y0 - 1 + ry, lz = lz0 - 1 + rz;
			const int z = lz - 1 + G.z0;
			if (y < 0 || y >= G.ny || lz < 0 || lz >= G.nlz || z < 0 || z >= G.nz) { continue; }
			const long long row = (long long)G.nx * (y + (long long)G.ny * lz);
			uint32_t pb = 0, pe = 0;
			if (xin) {
				pb = begin[row + x];
				pe = begin[row + x + 1];
			}
			const int cnt = (int)(pe - pb);
			int maxcnt = cnt;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o)); }
			if (maxcnt == 0) { continue; }
			const double yc = cys[y], zc = czs[z];
			cc[3] = y > 0 ? cys[y - 1] : yc - G.h;
			cc[4] = yc;
			cc[5] = y + 1 < G.ny ? cys[y + 1] : yc + G.h;
			cc[6] = z > 0 ? czs[z - 1] : zc - G.h;
			cc[7] = zc;
			cc[8] = z + 1 < G.nz ? czs[z + 1] : zc + G.h;
			row_component<COMP, APIC>(st, acc, fields, permv, permc, lane, pb, pe, cnt, maxcnt, cc, Q, ry, rz, nfx);
		}
		__syncthreads(); // rows of the next colour may touch the faces this colour just updated
	}
}

template <int METHOD> __global__ void __launch_bounds__(PB_THREADS, 2) k_p2g_brick(GridDesc G, PBParams Q,
	ParticleSoA P, const uint32_t *__restrict__ permv, const uint32_t *__restrict__ permc,
	const uint32_t *__restrict__ begin, const double *__restrict__ cxs,
	const double *__restrict__ cys, const double *__restrict__ czs, double *__restrict__ u, double *__restrict__ v,
	double *__restrict__ w, double *__restrict__ uo, double *__restrict__ vo, double *__restrict__ wo,
	uint8_t *__restrict__ typ) {
	constexpr bool APIC = METHOD == LFK_METHOD_APIC;
	extern __shared__ double smem[];
	double *acc = smem;                              // [2][BZ][BY][32], reused by the three components
	double *stage_all = smem + PB_ACC_COMP;          // [WARPS][FIELDS][32 * CSTRIDE]
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	double *st = stage_all + warp * (PB_FIELDS * PB_FSTRIDE);
	const int x0 = blockIdx.x * PB_BX, y0 = blockIdx.y * PB_BY, lz0 = blockIdx.z * PB_BZ + 1;
	const int nfx = min(PB_BX, G.nx - x0); // faces of this brick along x
	const double *fields[15];
#pragma unroll
	for (int f = 0; f < 15; ++f) { fields[f] = P.f[f]; }
	double *const out[3] = { u, v, w };
	double *const old[3] = { uo, vo, wo };

#pragma unroll 1
	for (int comp = 0; comp < 3; ++comp) {
		for (int e = threadIdx.x; e < PB_ACC_COMP; e += PB_THREADS) { acc[e] = 0.0; }
		__syncthreads();
		if (comp == 0) {
			brick_component<0, APIC>(G, Q, st, acc, fields, permv, permc, begin, cxs, cys, czs, warp, lane, x0, y0, lz0, nfx);
		} else if (comp == 1) {
			brick_component<1, APIC>(G, Q, st, acc, fields, permv, permc, begin, cxs, cys, czs, warp, lane, x0, y0, lz0, nfx);
		} else {
			brick_component<2, APIC>(G, Q, st, acc, fields, permv, permc, begin, cxs, cys, czs, warp, lane, x0, y0, lz0, nfx);
		}
		// ---- normalise + boundary faces + FLIP snapshot + gravity, write this component of the brick's faces
		// (the last colour ended with a __syncthreads) ----
		for (int e = threadIdx.x; e < PB_BZ * PB_BY * 32; e += PB_THREADS) {
			const int fx = e & 31, fy = (e >> 5) % PB_BY, fz = (e >> 5) / PB_BY;
			const int cx = x0 + fx, cyy = y0 + fy, lz = lz0 + fz;
			if (fx >= nfx || cyy >= G.ny || lz > G.nzl) { continue; }
			const int z = lz - 1 + G.z0;
			const long long me = cx + (long long)G.nx * (cyy + (long long)G.ny * lz);
			const double sw = acc[e], sv = acc[PB_BZ * PB_BY * 32 + e];
			double r = sw > 1e-6 ? sv / sw : 0.0; // src/simulation.cpp:380-386
			const bool edge = comp == 0 ? cx == G.nx - 1 : (comp == 1 ? cyy == G.ny - 1 : z == G.nz - 1);
			if (METHOD == LFK_METHOD_FLIP) { old[comp][me] = edge ? 0.0 : r; } // :340-344
			if (APIC && edge) { r = 0.0; }                                      // :397
			if (Q.add_gravity) { r += Q.gdt[comp]; }                            // :72-78
			out[comp][me] = r;
			if (comp == 0) { // classification, once per cell (:388-393)
				uint8_t t = typ[me];
				if (t != LFK_CELL_SOLID) {
					typ[me] = (begin[me + 1] - begin[me]) > 0 ? LFK_CELL_FLUID : LFK_CELL_AIR;
				}
			}
		}
		__syncthreads();
	}
}

int lfkg_p2g_brick(lfk_ctx *c, double gravity_dt, bool add_gravity) {
	const GridDesc &G = c->g;
	PBParams Q;
	Q.half = 0.5 * G.h;
	Q.inv_h = 1.0 / G.h;
	Q.add_gravity = add_gravity ? 1 : 0;
	for (int d = 0; d < 3; ++d) {
		Q.gdt[d] = c->prm.gravity[d] * gravity_dt;
	}
	dim3 grid((unsigned)((G.nx + PB_BX - 1) / PB_BX), (unsigned)((G.ny + PB_BY - 1) / PB_BY),
		(unsigned)((G.nzl + PB_BZ - 1) / PB_BZ));
	const uint32_t *permv = c->v_deferred ? c->perm : nullptr, *permc = c->c_deferred ? c->perm : nullptr;
	const size_t smem = (size_t)(PB_ACC_COMP + PB_WARPS * PB_FIELDS * PB_FSTRIDE) * sizeof(double);
	static bool attr_set[LFK_MAX_DEVICES] = {}; // function attributes are per device
	if (!attr_set[c->device % LFK_MAX_DEVICES]) {
		LFK_CUDA(c, cudaFuncSetAttribute(k_p2g_brick<LFK_METHOD_PIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LFK_CUDA(c, cudaFuncSetAttribute(k_p2g_brick<LFK_METHOD_FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LFK_CUDA(c, cudaFuncSetAttribute(k_p2g_brick<LFK_METHOD_APIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set[c->device % LFK_MAX_DEVICES] = true;
	}
	switch (c->prm.method) {
	case LFK_METHOD_PIC:
		LFK_LAUNCH(c, k_p2g_brick<LFK_METHOD_PIC>, grid, PB_THREADS, smem, G, Q, c->P, permv, permc, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	case LFK_METHOD_FLIP:
		LFK_LAUNCH(c, k_p2g_brick<LFK_METHOD_FLIP>, grid, PB_THREADS, smem, G, Q, c->P, permv, permc, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	default:
		LFK_LAUNCH(c, k_p2g_brick<LFK_METHOD_APIC>, grid, PB_THREADS, smem, G, Q, c->P, permv, permc, c->begin, c->ctr[0], c->ctr[1],
			c->ctr[2], c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	}
	return 0;
}
