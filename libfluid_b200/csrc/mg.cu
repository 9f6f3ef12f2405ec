// GPU-parallel preconditioner that replaces the reference's sequential MIC(0) (src/pressure_solver.cpp:244-332):
// one aggregation-multigrid V-cycle per PCG iteration.
//
//   * levels are dense grids, each 2x2x2 coarsening of the one below (piecewise-constant prolongation P);
//   * coarse operators are GALERKIN products P^T A P.  With 2x2x2 aggregates of a 7-point operator the product is
//     again a 7-point operator: a diagonal and one coupling per +face, all small non-negative integers (exact in
//     fp32).  Solid walls (no coupling) and the free surface (Dirichlet: diagonal counts the air neighbour) are
//     represented exactly on every level, which is what makes this robust on irregular fluid domains;
//   * smoother: red-black Gauss-Seidel, V(2,2), red-black before / black-red after the coarse correction, so the
//     cycle is a symmetric positive definite operator (required by CG);
//   * the coarse correction is over-relaxed (x += omega * P e, omega = 1.8): piecewise-constant aggregation
//     under-estimates the correction by about 2x in 3-D; tools/mg_prototype.py measures 12-14 PCG iterations at
//     128^3 with omega in [1.8, 2] against 19+ with omega = 1 and ~230+ for Jacobi;
//   * the whole cycle runs in fp32 (the preconditioner only shapes the search directions; CG itself stays fp64),
//     which halves its memory traffic.
//
// Pass fusion (every pass over level 0 costs ~13 B/cell of HBM traffic, so passes are what is minimised):
//   * the first red half-sweep from the zero initial guess is produced by the PCG kernel that writes r (MgPreload);
//   * the prolongation is folded into the first post-smoothing half-sweep: a Gauss-Seidel update overwrites its
//     cell without reading it, so `x += omega P e` only matters for the RED neighbours that the first (black)
//     post-sweep reads -- it adds omega * e_coarse to them on the fly -- and the red cells themselves are
//     overwritten by the red sweep that follows.  No prolongation pass exists;
//   * the last (red) half-sweep also converts the result to fp64 z, and accumulates z.r (k_mg_final_l0).
//
// Level 0 reads a 2-byte-per-cell coupling mask (built from the solver's flag byte once per solve) instead of
// coefficient arrays.
#include "lfk_internal.cuh"

#include <algorithm>

// ---- storage of the level-0 vectors: fp32 (fp16 storage was measured in r2a and changed nothing: 1.039 against
// 1.035 ms per iteration at 256^3 -- the level-0 passes are not limited by their bytes; removed) --------------------
__device__ __forceinline__ float l0_ld_b(const float *__restrict__ p, long long i) { return p[i]; }
__device__ __forceinline__ float l0_ld_x(const float *__restrict__ p, long long i) { return p[i]; }
__device__ __forceinline__ void l0_st_x(float *__restrict__ p, long long i, float v) { p[i] = v; }

#ifndef L0_U
#define L0_U 4  // cells in flight per thread in the level-0 half-sweeps (rows_pipelined)
#endif
#ifndef FIN_U
#define FIN_U 4 // ... and in the last half-sweep (fp64 z, z.r)
#endif
#define MG_OMEGA 1.8f
#define MG_PRE 2
#define MG_POST 2
#define MG_COARSE_SWEEPS 8
#define MG_COARSE_SWEEPS_MULTI 4 // multi-GPU coarsest level (host-driven sweeps, one NCCL exchange per half-sweep)
#define MG_COARSE_MAX_CELLS 4096 // a level this small is smoothed to convergence by one block

// level-0 coupling mask: bits 0-2 diagonal (non-solid neighbour count), bit 3 "is an unknown", bits 4-9 "the
// -x, +x, -y, +y, -z, +z neighbour is an unknown coupled to this cell"
#define MF_N(m) ((m) & 7u)
#define MF_L 8u
#define MF_XM 16u
#define MF_XP 32u
#define MF_YM 64u
#define MF_YP 128u
#define MF_ZM 256u
#define MF_ZP 512u

struct LevelDev { // by-value kernel argument
	int nx, ny, nzl;   // owned size
	int zpar;          // parity offset of the first owned layer (global z0 of this level)
	long long sxy, nown;
	const float *diag, *cx, *cy, *cz;
	float *x, *b;
};

static LevelDev level_dev(const MgLevel &L, int z0) {
	LevelDev d;
	d.nx = L.nx; d.ny = L.ny; d.nzl = L.nzl; d.zpar = z0 & 1;
	d.sxy = L.sxy; d.nown = L.sxy * L.nzl;
	d.diag = L.diag; d.cx = L.cx; d.cy = L.cy; d.cz = L.cz; d.x = L.x; d.b = L.b;
	return d;
}

// warp-per-row iteration over the owned cells of a level (no 64-bit div/mod); lanes <-> consecutive x
template <typename F> __device__ __forceinline__ void for_rows(int nx, int ny, int nzl, F f) {
	const int wpb = (int)(blockDim.x >> 5), rows = ny * nzl;
	for (int row = (int)blockIdx.x * wpb + (int)(threadIdx.x >> 5); row < rows; row += (int)gridDim.x * wpb) {
		const int y = row % ny, lz = row / ny + 1;
		const long long base = (long long)nx * (y + (long long)ny * lz);
		for (int x = (int)(threadIdx.x & 31); x < nx; x += 32) {
			f(x, y, lz, base + x);
		}
	}
}
// same, but only the cells of one colour: lane <-> every second cell of the row
template <typename F> __device__ __forceinline__ void for_rows_colour(int nx, int ny, int nzl, int zpar, int colour, F f) {
	const int wpb = (int)(blockDim.x >> 5), rows = ny * nzl;
	for (int row = (int)blockIdx.x * wpb + (int)(threadIdx.x >> 5); row < rows; row += (int)gridDim.x * wpb) {
		const int y = row % ny, lz = row / ny + 1;
		const long long base = (long long)nx * (y + (long long)ny * lz);
		for (int x = 2 * (int)(threadIdx.x & 31) + ((y + (lz - 1 + zpar) + colour) & 1); x < nx; x += 64) {
			f(x, y, lz, base + x);
		}
	}
}
static inline unsigned row_blocks(int ny, int nzl, int threads, unsigned cap) {
	long long rows = (long long)ny * nzl, wpb = threads / 32;
	long long nb = (rows + wpb - 1) / wpb;
	if (nb < 1) { nb = 1; }
	return (unsigned)(nb > cap ? cap : nb);
}

// ---- level 0 ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mg_mask_l0(GridDesc G, const uint8_t *__restrict__ flags,
	uint16_t *__restrict__ mask) {
	for_own_cells(G, [&](int x, int y, int, long long c) {
		unsigned f = flags[c], m = 0;
		if (f & FL_L) {
			m = FL_N(f) | MF_L;
			if (f & FL_SELF) { // coupled to the -neighbours that are unknowns (their "+ neighbour is fluid" == this cell)
				if (x > 0 && (flags[c - 1] & FL_L)) { m |= MF_XM; }
				if (y > 0 && (flags[c - G.nx] & FL_L)) { m |= MF_YM; }
				if (flags[c - G.sxy] & FL_L) { m |= MF_ZM; }
			}
			if ((f & FL_XP) && (flags[c + 1] & FL_L)) { m |= MF_XP; }
			if ((f & FL_YP) && (flags[c + G.nx] & FL_L)) { m |= MF_YP; }
			if ((f & FL_ZP) && (flags[c + G.sxy] & FL_L)) { m |= MF_ZP; }
		}
		mask[c] = (uint16_t)m;
	});
}

// All six neighbour loads are issued unconditionally (the arrays carry a ghost layer in z and rows are contiguous,
// so every address is inside the allocation) and masked afterwards: one memory round trip per cell instead of
// "mask, then the neighbours the mask selects".
struct L0Raw {
	float b, xc, xm, xp, ym, yp, zm, zp;
	float e, ex, ey, ez; // coarse corrections (PROLONG only)
	unsigned m;
};
template <bool PROLONG, typename T> __device__ __forceinline__ L0Raw l0_load(const GridDesc &G,
	const uint16_t *__restrict__ mask, const T *__restrict__ b, const T *__restrict__ X, const LevelDev &C, int x, int y,
	int lz, long long c) {
	L0Raw r;
	r.m = mask[c];
	r.b = l0_ld_b(b, c);
	r.xc = l0_ld_x(X, c);
	r.xm = l0_ld_x(X, c - 1);
	r.xp = l0_ld_x(X, c + 1);
	r.ym = l0_ld_x(X, c - G.nx);
	r.yp = l0_ld_x(X, c + G.nx);
	r.zm = l0_ld_x(X, c - G.sxy);
	r.zp = l0_ld_x(X, c + G.sxy);
	if (PROLONG) {
		const int X_ = x >> 1, Y_ = y >> 1, LZ = ((lz - 1) >> 1) + 1;
		const long long cc = X_ + (long long)C.nx * (Y_ + (long long)C.ny * LZ);
		// the neighbour across the aggregate boundary belongs to the adjacent aggregate, the other one to this one
		r.e = C.x[cc];
		r.ex = C.x[cc + ((x & 1) ? 1 : -1)];
		r.ey = C.x[cc + ((y & 1) ? C.nx : -C.nx)];
		r.ez = C.x[cc + (((lz - 1) & 1) ? C.sxy : -C.sxy)];
	}
	return r;
}
__device__ __forceinline__ float l0_offdiag_sum(const L0Raw &r) {
	const unsigned m = r.m;
	float s = 0.f;
	s += (m & MF_XM) ? r.xm : 0.f;
	s += (m & MF_XP) ? r.xp : 0.f;
	s += (m & MF_YM) ? r.ym : 0.f;
	s += (m & MF_YP) ? r.yp : 0.f;
	s += (m & MF_ZM) ? r.zm : 0.f;
	s += (m & MF_ZP) ? r.zp : 0.f;
	return s;
}
// omega * (sum over the coupled neighbours of the coarse correction of THEIR aggregate)
__device__ __forceinline__ float l0_prolong_sum(const L0Raw &r, int x, int y, int lz) {
	const unsigned m = r.m;
	float s = 0.f;
	s += (m & MF_XM) ? ((x & 1) ? r.e : r.ex) : 0.f;
	s += (m & MF_XP) ? ((x & 1) ? r.ex : r.e) : 0.f;
	s += (m & MF_YM) ? ((y & 1) ? r.e : r.ey) : 0.f;
	s += (m & MF_YP) ? ((y & 1) ? r.ey : r.e) : 0.f;
	s += (m & MF_ZM) ? (((lz - 1) & 1) ? r.e : r.ez) : 0.f;
	s += (m & MF_ZP) ? (((lz - 1) & 1) ? r.ez : r.e) : 0.f;
	return MG_OMEGA * s;
}

// one colour of red-black Gauss-Seidel on level 0.  PROLONG: the neighbours carry a pending coarse correction.
template <bool PROLONG, typename T> __global__ void __launch_bounds__(256) k_mg_rbgs_l0(GridDesc G,
	const uint16_t *__restrict__ mask, const T *__restrict__ b, T *__restrict__ X, int colour, LevelDev C,
	const PcgScalars *scal) {
	if (scal->done) { return; }
	rows_pipelined<L0_U, L0Raw>(G.nx, G.ny, G.nzl, G.z0, colour,
		[&](int x, int y, int lz, long long c) { return l0_load<PROLONG, T>(G, mask, b, X, C, x, y, lz, c); },
		[&](int x, int y, int lz, long long c, const L0Raw &r) {
			float s = r.b + l0_offdiag_sum(r);
			if (PROLONG) { s += l0_prolong_sum(r, x, y, lz); }
			if ((r.m & MF_L) && MF_N(r.m) > 0) { l0_st_x(X, c, __fdividef(s, (float)MF_N(r.m))); }
		});
}

// last half-sweep of the cycle (colour `colour`) fused with z = x0 / a_scale (fp64), sigma_new = z.r and its finaliser
template <typename T> __global__ void __launch_bounds__(RED_THREADS) k_mg_final_l0(GridDesc G,
	const uint16_t *__restrict__ mask, const T *__restrict__ b, const T *__restrict__ X, int colour,
	const double *__restrict__ r, double *__restrict__ z, PcgScalars *scal, double *partials,
	unsigned *ticket, int finalize, int first) {
	if (scal->done) { return; }
	const double inv_a_scale = scal->inv_a_scale;
	double acc = 0.0;
	struct FinRaw { L0Raw l; double r; };
	LevelDev none{};
	rows_pipelined<FIN_U, FinRaw>(G.nx, G.ny, G.nzl, 0, -1,
		[&](int x, int y, int lz, long long c) {
			FinRaw v;
			v.l = l0_load<false, T>(G, mask, b, X, none, x, y, lz, c);
			v.r = r[c];
			return v;
		},
		[&](int x, int y, int lz, long long c, const FinRaw &v) {
			const unsigned m = v.l.m;
			const float upd = __fdividef(v.l.b + l0_offdiag_sum(v.l), (float)(MF_N(m) > 0 ? MF_N(m) : 1u));
			const bool mine = (((x + y + (lz - 1 + G.z0)) & 1) == colour) && MF_N(m) > 0;
			const double zv = (m & MF_L) ? (double)(mine ? upd : v.l.xc) * inv_a_scale : 0.0;
			acc += zv * v.r; // zv == 0 on cells that are not unknowns
			z[c] = zv;
		});
	acc = block_sum(acc);
	if (threadIdx.x == 0) { partials[blockIdx.x] = acc; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 0);
		if (threadIdx.x == 0) {
			scal->sigma_new = tot;
			if (finalize) { pcg_finalize(scal, first ? FIN_BETA_FIRST : FIN_BETA); }
		}
	}
}

// coarse b = P^T (b - A x) of level 0, coarse x = 0.  One thread per coarse cell.
template <typename T> __global__ void __launch_bounds__(128) k_mg_restrict_l0(GridDesc G,
	const uint16_t *__restrict__ mask, const T *__restrict__ b, const T *__restrict__ X, LevelDev C,
	const PcgScalars *scal) {
	if (scal->done) { return; }
	// the pre-smoothing ended with a black half-sweep: black residuals are zero (to rounding), so only the 4 red
	// cells of each aggregate are visited; (dy, dz) in {0,1}^2, dx fixed by the colour
	struct Raw4 { L0Raw q[4]; };
	LevelDev none{};
	rows_pipelined<1, Raw4>(C.nx, C.ny, C.nzl, 0, -1,
		[&](int X_, int Y_, int LZ, long long) {
			Raw4 v;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				int y = 2 * Y_ + (k & 1), lz = 2 * (LZ - 1) + ((k >> 1) & 1) + 1;
				int x = 2 * X_ + ((y + (lz - 1 + G.z0)) & 1);
				// cells beyond an odd-sized grid: clamp the address (any valid cell), the contribution is dropped below
				x = x < G.nx ? x : G.nx - 1;
				y = y < G.ny ? y : G.ny - 1;
				lz = lz <= G.nzl ? lz : G.nzl;
				v.q[k] = l0_load<false, T>(G, mask, b, X, none, x, y, lz, x + (long long)G.nx * (y + (long long)G.ny * lz));
			}
			return v;
		},
		[&](int X_, int Y_, int LZ, long long cc, const Raw4 &v) {
			float acc = 0.f;
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int y = 2 * Y_ + (k & 1), lz = 2 * (LZ - 1) + ((k >> 1) & 1) + 1;
				const int x = 2 * X_ + ((y + (lz - 1 + G.z0)) & 1);
				const L0Raw &q = v.q[k];
				const float res = q.b - ((float)MF_N(q.m) * q.xc - l0_offdiag_sum(q));
				acc += ((q.m & MF_L) && x < G.nx && y < G.ny && lz <= G.nzl) ? res : 0.f;
			}
			C.b[cc] = acc;
			C.x[cc] = 0.f;
		});
}

// ---- generic level: coefficient arrays ------------------------------------------------------------------------
__device__ __forceinline__ float lv_offdiag_sum(const LevelDev &L, const float *__restrict__ X, long long c) {
	// couplings across the domain boundary are 0, so the wrapped neighbour reads are harmless (0 * finite)
	return L.cx[c] * X[c + 1] + L.cx[c - 1] * X[c - 1] + L.cy[c] * X[c + L.nx] + L.cy[c - L.nx] * X[c - L.nx] +
		L.cz[c] * X[c + L.sxy] + L.cz[c - L.sxy] * X[c - L.sxy];
}
template <bool PROLONG> __global__ void __launch_bounds__(256) k_mg_rbgs(LevelDev L, int colour, LevelDev C,
	const PcgScalars *scal) {
	if (scal->done) { return; }
	struct Raw {
		float d, b, cxm, cxp, cym, cyp, czm, czp, xm, xp, ym, yp, zm, zp, e, ex, ey, ez;
	};
	rows_pipelined<2, Raw>(L.nx, L.ny, L.nzl, L.zpar, colour,
		[&](int x, int y, int lz, long long c) {
			Raw r;
			r.d = L.diag[c];
			r.b = L.b[c];
			// couplings across the domain boundary are 0, so the wrapped neighbour reads are harmless (0 * finite)
			r.cxm = L.cx[c - 1]; r.cxp = L.cx[c];
			r.cym = L.cy[c - L.nx]; r.cyp = L.cy[c];
			r.czm = L.cz[c - L.sxy]; r.czp = L.cz[c];
			r.xm = L.x[c - 1]; r.xp = L.x[c + 1];
			r.ym = L.x[c - L.nx]; r.yp = L.x[c + L.nx];
			r.zm = L.x[c - L.sxy]; r.zp = L.x[c + L.sxy];
			if (PROLONG) {
				const int X_ = x >> 1, Y_ = y >> 1, LZ = ((lz - 1) >> 1) + 1;
				const long long cc = X_ + (long long)C.nx * (Y_ + (long long)C.ny * LZ);
				r.e = C.x[cc];
				r.ex = C.x[cc + ((x & 1) ? 1 : -1)];
				r.ey = C.x[cc + ((y & 1) ? C.nx : -C.nx)];
				r.ez = C.x[cc + (((lz - 1) & 1) ? C.sxy : -C.sxy)];
			}
			return r;
		},
		[&](int x, int y, int lz, long long c, const Raw &r) {
			float s = r.b + (r.cxp * r.xp + r.cxm * r.xm + r.cyp * r.yp + r.cym * r.ym + r.czp * r.zp + r.czm * r.zm);
			if (PROLONG) {
				// a zero coupling multiplies a finite value: the coarse arrays carry a ghost shell of zeros
				float t = r.cxm * ((x & 1) ? r.e : r.ex) + r.cxp * ((x & 1) ? r.ex : r.e);
				t += r.cym * ((y & 1) ? r.e : r.ey) + r.cyp * ((y & 1) ? r.ey : r.e);
				t += r.czm * (((lz - 1) & 1) ? r.e : r.ez) + r.czp * (((lz - 1) & 1) ? r.ez : r.e);
				s += MG_OMEGA * t;
			}
			if (r.d > 0.f) { L.x[c] = __fdividef(s, r.d); }
		});
}

__global__ void __launch_bounds__(128) k_mg_restrict(LevelDev F, LevelDev C, const PcgScalars *scal) {
	if (scal->done) { return; }
	for_rows(C.nx, C.ny, C.nzl, [&](int X_, int Y_, int LZ, long long cc) {
		float acc = 0.f;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			int x = 2 * X_ + (k & 1), y = 2 * Y_ + ((k >> 1) & 1), lz = 2 * (LZ - 1) + ((k >> 2) & 1) + 1;
			if (x >= F.nx || y >= F.ny || lz > F.nzl) { continue; }
			long long c = x + (long long)F.nx * (y + (long long)F.ny * lz);
			float d = F.diag[c];
			if (d <= 0.f) { continue; }
			acc += F.b[c] - (d * F.x[c] - lv_offdiag_sum(F, F.x, c));
		}
		C.b[cc] = acc;
		C.x[cc] = 0.f;
	});
}

// ---- the coarse tail: every level with at most MG_COARSE_MAX_CELLS cells is handled by ONE block in ONE launch
// (pre-smooth / restrict down, symmetric sweeps on the coarsest grid, prolong / post-smooth up), which removes
// ~10 tiny launches per level from every PCG iteration.
#define MG_TAIL_MAX_LEVELS 8
struct TailLevels {
	LevelDev L[MG_TAIL_MAX_LEVELS];
	int n;
	int coarse_sweeps; // symmetric sweeps on the coarsest level
};

// block-wide loop over the owned cells of a small level (int arithmetic only)
template <typename F> __device__ __forceinline__ void tail_cells(const LevelDev &L, F f) {
	const int n = (int)L.nown;
	for (int own = (int)threadIdx.x; own < n; own += (int)blockDim.x) {
		const int x = own % L.nx, rest = own / L.nx;
		f(x, rest % L.ny, rest / L.ny + 1, (long long)own + L.sxy);
	}
}
__device__ __forceinline__ void tail_half_sweep(const LevelDev &L, int colour) {
	tail_cells(L, [&](int x, int y, int lz, long long c) {
		if (((x + y + (lz - 1 + L.zpar) + colour) & 1) != 0) { return; }
		float d = L.diag[c];
		if (d > 0.f) {
			L.x[c] = (L.b[c] + lv_offdiag_sum(L, L.x, c)) / d;
		}
	});
	__syncthreads();
}
__device__ __forceinline__ void tail_restrict(const LevelDev &F, const LevelDev &C) {
	tail_cells(C, [&](int X_, int Y_, int LZ, long long cc) {
		float acc = 0.f;
		for (int k = 0; k < 8; ++k) {
			int x = 2 * X_ + (k & 1), y = 2 * Y_ + ((k >> 1) & 1), lz = 2 * (LZ - 1) + ((k >> 2) & 1) + 1;
			if (x >= F.nx || y >= F.ny || lz > F.nzl) { continue; }
			long long c = x + (long long)F.nx * (y + (long long)F.ny * lz);
			float d = F.diag[c];
			if (d <= 0.f) { continue; }
			acc += F.b[c] - (d * F.x[c] - lv_offdiag_sum(F, F.x, c));
		}
		C.b[cc] = acc;
		C.x[cc] = 0.f;
	});
	__syncthreads();
}
__device__ __forceinline__ void tail_prolong(const LevelDev &F, const LevelDev &C) {
	tail_cells(F, [&](int x, int y, int lz, long long c) {
		if (F.diag[c] <= 0.f) { return; }
		long long cc = (x >> 1) + (long long)C.nx * ((y >> 1) + (long long)C.ny * (((lz - 1) >> 1) + 1));
		F.x[c] += MG_OMEGA * C.x[cc];
	});
	__syncthreads();
}

__device__ __forceinline__ void tail_cycle(const TailLevels &T);

__global__ void __launch_bounds__(1024) k_mg_tail(TailLevels T, const PcgScalars *scal) {
	if (scal->done) { return; }
	tail_cycle(T);
}

// Production variant: the tail levels (a few thousand cells in all) are copied into shared memory first, so the ~70
// dependent half-sweeps / transfers of the tail run at shared-memory latency instead of L2 latency.
__global__ void __launch_bounds__(1024) k_mg_tail_smem(TailLevels T, const PcgScalars *scal) {
	if (scal->done) { return; }
	extern __shared__ float tail_sm[];
	TailLevels S = T;
	float *ptr = tail_sm;
	for (int l = 0; l < T.n; ++l) {
		const LevelDev &L = T.L[l];
		const int n = (int)(L.sxy * (L.nzl + 2)) + 2;
		float *d = ptr, *cx = ptr + n, *cy = ptr + 2 * n, *cz = ptr + 3 * n, *x = ptr + 4 * n, *b = ptr + 5 * n;
		ptr += 6 * n;
		for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
			d[i] = L.diag[i];
			cx[i] = L.cx[i];
			cy[i] = L.cy[i];
			cz[i] = L.cz[i];
			x[i] = l == 0 ? L.x[i] : 0.f; // deeper levels are filled by the restrictions below; ghost entries stay 0
			b[i] = l == 0 ? L.b[i] : 0.f;
		}
		S.L[l].diag = d; S.L[l].cx = cx; S.L[l].cy = cy; S.L[l].cz = cz; S.L[l].x = x; S.L[l].b = b;
	}
	__syncthreads();
	tail_cycle(S);
	const LevelDev &L0 = T.L[0];
	const int n0 = (int)(L0.sxy * (L0.nzl + 2)) + 2;
	for (int i = (int)threadIdx.x; i < n0; i += (int)blockDim.x) { L0.x[i] = S.L[0].x[i]; }
}

__device__ __forceinline__ void tail_cycle(const TailLevels &T) {
	const int last = T.n - 1;
	for (int l = 0; l < last; ++l) { // down: x of the entry level was zeroed by the restriction that filled its b
		for (int s = 0; s < MG_PRE; ++s) {
			tail_half_sweep(T.L[l], 0);
			tail_half_sweep(T.L[l], 1);
		}
		tail_restrict(T.L[l], T.L[l + 1]);
	}
	for (int s = 0; s < T.coarse_sweeps; ++s) { // coarsest: red, black, black, red
		tail_half_sweep(T.L[last], 0);
		tail_half_sweep(T.L[last], 1);
		tail_half_sweep(T.L[last], 1);
		tail_half_sweep(T.L[last], 0);
	}
	for (int l = last - 1; l >= 0; --l) { // up
		tail_prolong(T.L[l], T.L[l + 1]);
		for (int s = 0; s < MG_POST; ++s) {
			tail_half_sweep(T.L[l], 1);
			tail_half_sweep(T.L[l], 0);
		}
	}
}

// ---- Galerkin coarse operators ----------------------------------------------------------------------------------
struct LevelOut {
	float *diag, *cx, *cy, *cz;
};
__global__ void __launch_bounds__(128) k_mg_build_l1(GridDesc G, const uint16_t *__restrict__ mask, LevelDev C,
	LevelOut O) {
	for_rows(C.nx, C.ny, C.nzl, [&](int X_, int Y_, int LZ, long long cc) {
		float diag = 0.f, cxp = 0.f, cyp = 0.f, czp = 0.f;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
			int x = 2 * X_ + dx, y = 2 * Y_ + dy, lz = 2 * (LZ - 1) + dz + 1;
			if (x >= G.nx || y >= G.ny || lz > G.nzl) { continue; }
			long long c = x + (long long)G.nx * (y + (long long)G.ny * lz);
			unsigned m = mask[c];
			if (!(m & MF_L)) { continue; }
			diag += (float)MF_N(m);
			// a coupling inside the aggregate folds into the diagonal (-1 from each side), one across it stays a coupling
			if (m & MF_XP) { if (dx == 0) { diag -= 2.f; } else { cxp += 1.f; } }
			if (m & MF_YP) { if (dy == 0) { diag -= 2.f; } else { cyp += 1.f; } }
			if (m & MF_ZP) {
				if (dz == 0 && lz + 1 <= G.nzl) { diag -= 2.f; } else { czp += 1.f; }
			}
		}
		O.diag[cc] = diag;
		O.cx[cc] = cxp;
		O.cy[cc] = cyp;
		O.cz[cc] = czp;
	});
}
__global__ void __launch_bounds__(128) k_mg_build(LevelDev F, LevelDev C, LevelOut O) {
	for_rows(C.nx, C.ny, C.nzl, [&](int X_, int Y_, int LZ, long long cc) {
		float diag = 0.f, cxp = 0.f, cyp = 0.f, czp = 0.f;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
			int x = 2 * X_ + dx, y = 2 * Y_ + dy, lz = 2 * (LZ - 1) + dz + 1;
			if (x >= F.nx || y >= F.ny || lz > F.nzl) { continue; }
			long long c = x + (long long)F.nx * (y + (long long)F.ny * lz);
			float d = F.diag[c];
			if (d <= 0.f) { continue; }
			diag += d;
			float a = F.cx[c], bq = F.cy[c], cq = F.cz[c];
			if (dx == 0) { diag -= 2.f * a; } else { cxp += a; }
			if (dy == 0) { diag -= 2.f * bq; } else { cyp += bq; }
			if (dz == 0 && lz + 1 <= F.nzl) { diag -= 2.f * cq; } else { czp += cq; }
		}
		O.diag[cc] = diag;
		O.cx[cc] = cxp;
		O.cy[cc] = cyp;
		O.cz[cc] = czp;
	});
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int mg_alloc(lfk_ctx *c) {
	if (!c->mg.empty()) { return 0; }
	const GridDesc &G = c->g;
	int nx = G.nx, ny = G.ny, nzl = G.nzl, z0 = G.z0, nz = G.nz;
	// the slab layout of every rank (same rule as lfk_create)
	std::vector<int> all_z0((size_t)c->nranks), all_nzl((size_t)c->nranks);
	for (int r = 0; r < c->nranks; ++r) {
		const int base = G.nz / c->nranks, rem = G.nz % c->nranks;
		all_z0[r] = r * base + (r < rem ? r : rem);
		all_nzl[r] = base + (r < rem ? 1 : 0);
	}
	LFK_CUDA(c, cudaMalloc((void**)&c->mg_mask, ((size_t)G.ncl + 2) * sizeof(uint16_t)));
	LFK_CUDA(c, cudaMemsetAsync(c->mg_mask, 0, ((size_t)G.ncl + 2) * sizeof(uint16_t), c->stream));
	for (int l = 0; l < 16; ++l) {
		MgLevel L{};
		L.nx = nx; L.ny = ny; L.nzl = nzl; L.nlz = nzl + 2;
		L.sxy = (long long)nx * ny;
		L.ncl = L.sxy * L.nlz;
		size_t n = (size_t)L.ncl + 2; // +2: the wrapped neighbour reads of the last owned cell stay in bounds
		float **arrs[] = { &L.x, &L.b, &L.diag, &L.cx, &L.cy, &L.cz };
		int first = 0, last = l == 0 ? 2 : 6;
		for (int k = first; k < last; ++k) {
			LFK_CUDA(c, cudaMalloc((void**)arrs[k], n * sizeof(float)));
			LFK_CUDA(c, cudaMemsetAsync(*arrs[k], 0, n * sizeof(float), c->stream));
		}
		c->mg.push_back(L);
		c->mg_z0.push_back(z0);
		// slabs must stay aligned to the aggregates; stop coarsening when they would not (multi-GPU).  The decision is
		// taken over the slabs of ALL ranks (the layout is a pure function of nz, nranks and the level), so every rank
		// builds the same number of levels and the halo exchanges inside the V-cycle pair up.
		bool aligned = true;
		if (c->nranks > 1) {
			aligned = nz % 2 == 0;
			for (int r = 0; r < c->nranks && aligned; ++r) {
				aligned = all_z0[r] % 2 == 0 && all_nzl[r] % 2 == 0 && all_nzl[r] >= 2;
			}
		}
		if (!aligned || (nx <= 2 && ny <= 2 && nzl <= 2)) { break; }
		nx = (nx + 1) / 2; ny = (ny + 1) / 2; nzl = (nzl + 1) / 2; z0 /= 2; nz = (nz + 1) / 2;
		for (int r = 0; r < c->nranks; ++r) { all_z0[r] /= 2; all_nzl[r] = (all_nzl[r] + 1) / 2; }
	}
	return 0;
}

int lfkm_free(lfk_ctx *c) {
	for (MgLevel &L : c->mg) {
		float *arrs[] = { L.x, L.b, L.diag, L.cx, L.cy, L.cz };
		for (float *p : arrs) {
			if (p) { cudaFree(p); }
		}
	}
	if (c->mg_mask) { cudaFree(c->mg_mask); c->mg_mask = nullptr; }
	for (MgLevel &L : c->mg_agg) {
		float *arrs[] = { L.x, L.b, L.diag, L.cx, L.cy, L.cz };
		for (float *p : arrs) {
			if (p) { cudaFree(p); }
		}
	}
	c->mg_agg.clear();
	c->mg_agg_level = -1;
	c->mg.clear();
	c->mg_z0.clear();
	return 0;
}

// ---- multi-GPU: agglomerated coarse levels (lfk_set_tuning("mg_agg", 0) keeps every level distributed) ------------
// A small distributed level is pure latency: every half-sweep is a halo exchange (~10 us) for microseconds of arithmetic,
// and the hierarchy has to stop where the slabs stop being aligned to the 2x2x2 aggregates.  Instead, the first level
// whose WHOLE-GRID size is at most agg_max_cells() is assembled on EVERY rank (operators once per solve, right-hand side
// once per cycle: each rank writes its own layers into a zeroed global array and one all-reduce adds them up -- exact,
// every element has one non-zero contributor), the hierarchy continues below it on the global grid down to a handful
// of cells, every rank runs that part of the cycle redundantly (vcycle_agg: the generic level kernels, then the
// single-block tail), and copies its own layers plus the two ghost layers of the result back.  Per cycle: one
// all-reduce instead of ~8 exchanges per distributed level below that point and 16 on the coarsest one.
// Largest global level that is agglomerated: every distributed level costs ~8 halo exchanges of ~10 us per cycle however
// small it is, while a level of a few hundred thousand cells costs ~9 launches of ~5 us when every rank sweeps all of
// it -- so levels up to this size are cheaper replicated than distributed (2 GPUs, 512 x 256 x 256: 1.68 ms per PCG
// iteration with the threshold at the single-block tail size of 4096 cells against 2.06 ms fully distributed, r2d).
static long long agg_max_cells(const lfk_ctx *c) {
	return c->tune.mg_agg_cells > 0 ? (long long)c->tune.mg_agg_cells : 600000ll;
}

static int mg_agg_alloc(lfk_ctx *c) {
	if (!c->mg_agg.empty() || c->mg_agg_level == -2) { return 0; }
	c->mg_agg_level = -2; // decided: none (unless found below)
	const GridDesc &G = c->g;
	int la = -1;
	for (size_t l = 1; l < c->mg.size(); ++l) { // aligned coarsening: the global depth of level l is nz >> l
		const long long gnz = G.nz >> l;
		if ((long long)c->mg[l].nx * c->mg[l].ny * gnz <= agg_max_cells(c)) { la = (int)l; break; }
	}
	if (la < 0) { return 0; }
	int nx = c->mg[la].nx, ny = c->mg[la].ny, nz = G.nz >> la;
	for (int k = 0; k < 16; ++k) {
		MgLevel L{};
		L.nx = nx; L.ny = ny; L.nzl = nz; L.nlz = nz + 2;
		L.sxy = (long long)nx * ny;
		L.ncl = L.sxy * L.nlz;
		const size_t n = (size_t)L.ncl + 2;
		float **arrs[] = { &L.x, &L.b, &L.diag, &L.cx, &L.cy, &L.cz };
		for (float **a : arrs) {
			LFK_CUDA(c, cudaMalloc((void**)a, n * sizeof(float)));
			LFK_CUDA(c, cudaMemsetAsync(*a, 0, n * sizeof(float), c->stream));
		}
		c->mg_agg.push_back(L);
		if (nx <= 2 && ny <= 2 && nz <= 2) { break; }
		nx = (nx + 1) / 2; ny = (ny + 1) / 2; nz = (nz + 1) / 2;
	}
	c->mg_agg_level = la;
	return 0;
}

// own layers of a distributed level-`la` array -> their place in the zeroed global array, then the sum over the ranks
static int mg_agg_gather(lfk_ctx *c, const MgLevel &D, int z0, const float *src, const MgLevel &Gl, float *dst) {
	const size_t n = (size_t)Gl.ncl + 2;
	LFK_CUDA(c, cudaMemsetAsync(dst, 0, n * sizeof(float), c->stream));
	LFK_CUDA(c, cudaMemcpyAsync(dst + (size_t)Gl.sxy * (1 + z0), src + (size_t)D.sxy, (size_t)D.sxy * D.nzl * sizeof(float),
		cudaMemcpyDeviceToDevice, c->stream));
	return lfkx_allreduce_sum_f32(c, dst, (int)n);
}

static int mg_agg_setup(lfk_ctx *c) { // after the distributed operators of this solve exist
	LFK_TRY(mg_agg_alloc(c));
	if (c->mg_agg_level < 0) { return 0; }
	const int la = c->mg_agg_level;
	const MgLevel &D = c->mg[la];
	MgLevel &Gl = c->mg_agg[0];
	const int z0 = c->mg_z0[la];
	LFK_TRY(mg_agg_gather(c, D, z0, D.diag, Gl, Gl.diag));
	LFK_TRY(mg_agg_gather(c, D, z0, D.cx, Gl, Gl.cx));
	LFK_TRY(mg_agg_gather(c, D, z0, D.cy, Gl, Gl.cy));
	LFK_TRY(mg_agg_gather(c, D, z0, D.cz, Gl, Gl.cz));
	for (size_t k = 1; k < c->mg_agg.size(); ++k) {
		MgLevel &C = c->mg_agg[k];
		LevelDev Cd = level_dev(C, 0);
		LevelOut O{ C.diag, C.cx, C.cy, C.cz };
		LFK_LAUNCH(c, k_mg_build, row_blocks(Cd.ny, Cd.nzl, 128, 1u << 20), 128, 0, level_dev(c->mg_agg[k - 1], 0), Cd, O);
	}
	return 0;
}

static int launch_tail(lfk_ctx *c, const TailLevels &T);

// V-cycle on the agglomerated (replicated, whole-grid) levels >= k: the generic level kernels without any exchange;
// levels that fit the single-block tail are finished by it in one launch.  x of level k is zero on entry.
static int vcycle_agg(lfk_ctx *c, size_t k) {
	std::vector<MgLevel> &Ls = c->mg_agg;
	const size_t last = Ls.size() - 1;
	const LevelDev Ld = level_dev(Ls[k], 0);
	if (k == last || (Ld.nown <= MG_COARSE_MAX_CELLS && last - k < MG_TAIL_MAX_LEVELS)) {
		TailLevels T;
		T.n = 0;
		T.coarse_sweeps = c->tune.mg_coarse > 0 ? c->tune.mg_coarse : MG_COARSE_SWEEPS;
		for (size_t j = k; j <= last && T.n < MG_TAIL_MAX_LEVELS; ++j) { T.L[T.n++] = level_dev(Ls[j], 0); }
		return launch_tail(c, T);
	}
	const LevelDev Cd = level_dev(Ls[k + 1], 0);
	const unsigned nb = row_blocks(Ld.ny, Ld.nzl, 256, 1u << 20);
	for (int s = 0; s < MG_PRE; ++s) { // red, black
		LFK_LAUNCH(c, k_mg_rbgs<false>, nb, 256, 0, Ld, 0, Ld, c->d_scal);
		LFK_LAUNCH(c, k_mg_rbgs<false>, nb, 256, 0, Ld, 1, Ld, c->d_scal);
	}
	LFK_LAUNCH(c, k_mg_restrict, row_blocks(Cd.ny, Cd.nzl, 128, 1u << 20), 128, 0, Ld, Cd, c->d_scal);
	LFK_TRY(vcycle_agg(c, k + 1));
	for (int s = 0; s < MG_POST; ++s) { // black (the first one applies the prolongation on the fly), red
		if (s == 0) {
			LFK_LAUNCH(c, k_mg_rbgs<true>, nb, 256, 0, Ld, 1, Cd, c->d_scal);
		} else {
			LFK_LAUNCH(c, k_mg_rbgs<false>, nb, 256, 0, Ld, 1, Ld, c->d_scal);
		}
		LFK_LAUNCH(c, k_mg_rbgs<false>, nb, 256, 0, Ld, 0, Ld, c->d_scal);
	}
	return 0;
}

// the cycle of distributed level `la` and everything below it, on the agglomerated grid
static int mg_agg_cycle(lfk_ctx *c) {
	const int la = c->mg_agg_level;
	MgLevel &D = c->mg[la];
	MgLevel &Gl = c->mg_agg[0];
	const int z0 = c->mg_z0[la];
	LFK_TRY(mg_agg_gather(c, D, z0, D.b, Gl, Gl.b));
	LFK_CUDA(c, cudaMemsetAsync(Gl.x, 0, ((size_t)Gl.ncl + 2) * sizeof(float), c->stream));
	LFK_TRY(vcycle_agg(c, 0));
	// my layers and the ghost layer either side: global layer index of local layer 0 is z0 - 1 + 1 = z0
	LFK_CUDA(c, cudaMemcpyAsync(D.x, Gl.x + (size_t)Gl.sxy * z0, (size_t)D.sxy * (D.nzl + 2) * sizeof(float),
		cudaMemcpyDeviceToDevice, c->stream));
	return 0;
}

int lfkm_setup(lfk_ctx *c, double a_scale) {
	(void)a_scale;
	LFK_TRY(mg_alloc(c));
	const GridDesc &G = c->g;
	// the flag byte's ghost layers are valid here (lfks_build_system exchanged them)
	LFK_LAUNCH(c, k_mg_mask_l0, lfk_row_blocks(G, 256, 1u << 20), 256, 0, G, c->flags, c->mg_mask);
	for (size_t l = 1; l < c->mg.size(); ++l) {
		MgLevel &C = c->mg[l];
		LevelDev Cd = level_dev(C, c->mg_z0[l]);
		LevelOut O{ C.diag, C.cx, C.cy, C.cz };
		unsigned nb = row_blocks(Cd.ny, Cd.nzl, 128, 1u << 20);
		if (l == 1) {
			LFK_LAUNCH(c, k_mg_build_l1, nb, 128, 0, G, c->mg_mask, Cd, O);
		} else {
			LFK_LAUNCH(c, k_mg_build, nb, 128, 0, level_dev(c->mg[l - 1], c->mg_z0[l - 1]), Cd, O);
		}
		if (c->nranks > 1) { // -z couplings and the neighbour's diagonal live in the ghost layers
			LFK_TRY(lfkx_halo_f32(c, C.cz, C.nx, C.ny, C.nzl));
			LFK_TRY(lfkx_halo_f32(c, C.diag, C.nx, C.ny, C.nzl));
		}
	}
	if (c->nranks > 1 && c->tune.mg_agg == 1) { LFK_TRY(mg_agg_setup(c)); }
	c->mg_valid = true;
	return 0;
}

// one half-sweep of colour `colour` on level l; `prolong`: the neighbours carry the pending correction of level l + 1.
// Multi-GPU: the z ghost layers of x are refreshed first unless the caller knows they are current (`x_current`):
// every exchange is an NCCL launch of ~15 us against ~8 us of arithmetic on the coarse levels (profiles/r1d: 2.4 ms per
// PCG iteration on 2 GPUs against 1.03 ms on one), so the ones that cannot change anything are skipped:
//   * first half-sweep after a restriction: x == 0 on every rank, the ghost layers are zeroed locally instead;
//   * first post-smoothing half-sweep: x of this level has not changed since the exchange before the restriction.
static int half_sweep(lfk_ctx *c, size_t l, int colour, bool prolong, bool x_current = false) {
	const GridDesc &G = c->g;
	MgLevel &L = c->mg[l];
	LevelDev Ld = level_dev(L, c->mg_z0[l]);
	LevelDev Cd = prolong ? level_dev(c->mg[l + 1], c->mg_z0[l + 1]) : Ld;
	unsigned nb = row_blocks(L.ny, L.nzl, 256, 1u << 20);
	if (c->nranks > 1) {
		if (!x_current) { LFK_TRY(lfkx_halo_f32(c, L.x, L.nx, L.ny, L.nzl)); }
		if (prolong) { LFK_TRY(lfkx_halo_f32(c, c->mg[l + 1].x, Cd.nx, Cd.ny, Cd.nzl)); }
	}
	if (l == 0) {
		if (prolong) {
			LFK_LAUNCH(c, (k_mg_rbgs_l0<true, float>), nb, 256, 0, G, c->mg_mask, L.b, L.x, colour, Cd, c->d_scal);
		} else {
			LFK_LAUNCH(c, (k_mg_rbgs_l0<false, float>), nb, 256, 0, G, c->mg_mask, L.b, L.x, colour, Cd, c->d_scal);
		}
	} else {
		if (prolong) {
			LFK_LAUNCH(c, k_mg_rbgs<true>, nb, 256, 0, Ld, colour, Cd, c->d_scal);
		} else {
			LFK_LAUNCH(c, k_mg_rbgs<false>, nb, 256, 0, Ld, colour, Cd, c->d_scal);
		}
	}
	return 0;
}

// V-cycle on levels >= l.  At level 0 the first red half-sweep was pre-applied by the kernel that produced b0 / x0,
// and the very last (red) half-sweep is left to k_mg_final_l0.
// every level of T in one block, one launch (shared-memory copy of the levels when they fit)
static int launch_tail(lfk_ctx *c, const TailLevels &T) {
	size_t smem = 0;
	for (int k = 0; k < T.n; ++k) { smem += 6 * ((size_t)(T.L[k].sxy * (T.L[k].nzl + 2)) + 2) * sizeof(float); }
	if (smem <= 200 * 1024) {
		static bool attr_set[LFK_MAX_DEVICES] = {}; // function attributes are per device
		if (!attr_set[c->device % LFK_MAX_DEVICES]) {
			LFK_CUDA(c, cudaFuncSetAttribute(k_mg_tail_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
			attr_set[c->device % LFK_MAX_DEVICES] = true;
		}
		LFK_LAUNCH(c, k_mg_tail_smem, 1, 1024, smem, T, c->d_scal);
	} else {
		LFK_LAUNCH(c, k_mg_tail, 1, 1024, 0, T, c->d_scal);
	}
	return 0;
}

static int vcycle(lfk_ctx *c, size_t l) {
	const GridDesc &G = c->g;
	size_t last = c->mg.size() - 1;
	MgLevel &L = c->mg[l];
	LevelDev Ld = level_dev(L, c->mg_z0[l]);
	if (c->nranks == 1 && l > 0 && Ld.nown <= MG_COARSE_MAX_CELLS && last - l < MG_TAIL_MAX_LEVELS) {
		TailLevels T; // this level and everything below it: one block, one launch
		T.n = 0;
		T.coarse_sweeps = c->tune.mg_coarse > 0 ? c->tune.mg_coarse : MG_COARSE_SWEEPS;
		for (size_t k = l; k <= last; ++k) {
			T.L[T.n++] = level_dev(c->mg[k], c->mg_z0[k]);
		}
		return launch_tail(c, T);
	}
	if (c->nranks > 1 && c->tune.mg_agg == 1 && c->mg_agg_level > 0 && (int)l == c->mg_agg_level) {
		return mg_agg_cycle(c);
	}
	// x of a level > 0 is zero when its cycle starts (the restriction zeroes the owned cells, and -- multi-GPU -- the
	// ghost layers, see below), so its first half-sweep needs no exchange
	const bool zero_start = l > 0;
	if (l == last) { // coarsest level reached outside the tail (multi-GPU alignment limit, or a tiny fine grid)
		// multi-GPU, below level 0: a coarsest grid of a few cells per rank; tools/mg_prototype.py counts the same PCG
		// iterations with 4 symmetric sweeps as with 8 on slab hierarchies of 1 .. 8 ranks, and every half-sweep is an
		// exchange
		const int sweeps = (c->nranks > 1 && l > 0) ? MG_COARSE_SWEEPS_MULTI : MG_COARSE_SWEEPS;
		for (int s = 0; s < sweeps; ++s) {
			if (!(l == 0 && s == 0)) { LFK_TRY(half_sweep(c, l, 0, false, zero_start && s == 0)); }
			LFK_TRY(half_sweep(c, l, 1, false));
		}
		for (int s = 0; s < sweeps; ++s) {
			LFK_TRY(half_sweep(c, l, 1, false));
			if (!(l == 0 && s == sweeps - 1)) { LFK_TRY(half_sweep(c, l, 0, false)); }
		}
		return 0;
	}
	for (int s = 0; s < MG_PRE; ++s) { // red, black
		if (!(l == 0 && s == 0)) { LFK_TRY(half_sweep(c, l, 0, false, zero_start && s == 0)); }
		LFK_TRY(half_sweep(c, l, 1, false));
	}
	MgLevel &C = c->mg[l + 1];
	LevelDev Cd = level_dev(C, c->mg_z0[l + 1]);
	if (c->nranks > 1) { LFK_TRY(lfkx_halo_f32(c, L.x, L.nx, L.ny, L.nzl)); }
	unsigned rb = row_blocks(Cd.ny, Cd.nzl, 128, 1u << 20);
	if (l == 0) {
		LFK_LAUNCH(c, k_mg_restrict_l0<float>, rb, 128, 0, G, c->mg_mask, L.b, L.x, Cd, c->d_scal);
	} else {
		LFK_LAUNCH(c, k_mg_restrict, rb, 128, 0, Ld, Cd, c->d_scal);
	}
	if (c->nranks > 1) { // the coarse x starts from zero in the ghost layers too (what an exchange would deliver)
		LFK_CUDA(c, cudaMemsetAsync(C.x, 0, (size_t)C.sxy * sizeof(float), c->stream));
		LFK_CUDA(c, cudaMemsetAsync(C.x + (size_t)C.sxy * (C.nzl + 1), 0, (size_t)C.sxy * sizeof(float), c->stream));
	}
	LFK_TRY(vcycle(c, l + 1));
	for (int s = 0; s < MG_POST; ++s) { // black (the first one applies the prolongation on the fly), red
		// s == 0: x of this level is as the exchange before the restriction left it
		LFK_TRY(half_sweep(c, l, 1, s == 0, s == 0));
		if (!(l == 0 && s == MG_POST - 1)) { LFK_TRY(half_sweep(c, l, 0, false)); }
	}
	return 0;
}

// level-0 buffers, filled by the PCG kernels that produce r (see MgPreload in pressure.cu)
int lfkm_level0(lfk_ctx *c, float **b0, float **x0) {
	LFK_TRY(mg_alloc(c));
	*b0 = c->mg[0].b;
	*x0 = c->mg[0].x;
	return 0;
}

// z = M^-1 r, sigma_new = z.r.  b0 and the red half of x0 were written by k_pcg_init / k_update_pr.
int lfkm_apply_preloaded(lfk_ctx *c, unsigned nb, int fin, int first) {
	const GridDesc &G = c->g;
	MgLevel &L0 = c->mg[0];
	LFK_TRY(vcycle(c, 0));
	if (c->nranks > 1) { LFK_TRY(lfkx_halo_f32(c, L0.x, L0.nx, L0.ny, L0.nzl)); }
	{
		LFK_LAUNCH(c, k_mg_final_l0<float>, nb, RED_THREADS, 0, G, c->mg_mask, L0.b, L0.x, 0, c->r, c->z,
			c->d_scal, c->partials, c->ticket, fin, first);
	}
	return 0;
}
