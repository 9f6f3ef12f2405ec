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
// Level 0 reads the solver's flag byte (1 B/cell) instead of coefficient arrays.
#include "lfk_internal.cuh"

#include <algorithm>


#define MG_OMEGA 1.8f
#define MG_PRE 2
#define MG_POST 2
#define MG_COARSE_SWEEPS 8
#define MG_COARSE_MAX_CELLS 4096 // a level this small is smoothed to convergence by one block

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

// ---- level-0 operator from the flag byte ----------------------------------------------------------------------
__device__ __forceinline__ float l0_offdiag_sum(const GridDesc &G, unsigned f, const float *__restrict__ X,
	long long c, int x, int y) {
	float s = 0.f;
	if (f & FL_SELF) {
		if (x > 0) { s += X[c - 1]; }
		if (y > 0) { s += X[c - G.nx]; }
		s += X[c - G.sxy];
	}
	if (f & FL_XP) { s += X[c + 1]; }
	if (f & FL_YP) { s += X[c + G.nx]; }
	if (f & FL_ZP) { s += X[c + G.sxy]; }
	return s;
}

// one colour of red-black Gauss-Seidel on level 0; each thread owns one cell of that colour (x = 2i + parity)
__global__ void __launch_bounds__(256) k_mg_rbgs_l0(GridDesc G, const uint8_t *__restrict__ flags,
	const float *__restrict__ b, float *__restrict__ X, int colour, const PcgScalars *scal) {
	if (scal->done) { return; }
	int hx = (G.nx + 1) >> 1;
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long total = (long long)hx * G.ny * G.nzl;
	if (t >= total) { return; }
	int xi = (int)(t % hx);
	long long rest = t / hx;
	int y = (int)(rest % G.ny);
	int lz = (int)(rest / G.ny) + 1;
	int x = 2 * xi + ((y + (lz - 1 + G.z0) + colour) & 1);
	if (x >= G.nx) { return; }
	long long c = x + (long long)G.nx * (y + (long long)G.ny * lz);
	unsigned f = flags[c];
	if (!(f & FL_L) || FL_N(f) == 0) { return; }
	float s = b[c] + l0_offdiag_sum(G, f, X, c, x, y);
	X[c] = s / (float)FL_N(f);
}

// ---- generic level: coefficient arrays ------------------------------------------------------------------------
__device__ __forceinline__ float lv_offdiag_sum(const LevelDev &L, const float *__restrict__ X, long long c) {
	// couplings across the domain boundary are 0, so the wrapped neighbour reads are harmless (0 * finite)
	return L.cx[c] * X[c + 1] + L.cx[c - 1] * X[c - 1] + L.cy[c] * X[c + L.nx] + L.cy[c - L.nx] * X[c - L.nx] +
		L.cz[c] * X[c + L.sxy] + L.cz[c - L.sxy] * X[c - L.sxy];
}

__global__ void __launch_bounds__(256) k_mg_rbgs(LevelDev L, int colour, const PcgScalars *scal) {
	if (scal->done) { return; }
	int hx = (L.nx + 1) >> 1;
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long total = (long long)hx * L.ny * L.nzl;
	if (t >= total) { return; }
	int xi = (int)(t % hx);
	long long rest = t / hx;
	int y = (int)(rest % L.ny);
	int lz = (int)(rest / L.ny) + 1;
	int x = 2 * xi + ((y + (lz - 1 + L.zpar) + colour) & 1);
	if (x >= L.nx) { return; }
	long long c = x + (long long)L.nx * (y + (long long)L.ny * lz);
	float d = L.diag[c];
	if (d <= 0.f) { return; }
	L.x[c] = (L.b[c] + lv_offdiag_sum(L, L.x, c)) / d;
}

// ---- the coarse tail: every level with at most MG_COARSE_MAX_CELLS cells is handled by ONE block in ONE launch
// (pre-smooth / restrict down, symmetric sweeps on the coarsest grid, prolong / post-smooth up), which removes
// ~10 tiny launches per level from every PCG iteration.
#define MG_TAIL_MAX_LEVELS 8
struct TailLevels {
	LevelDev L[MG_TAIL_MAX_LEVELS];
	int n;
};

__device__ __forceinline__ void tail_half_sweep(const LevelDev &L, int colour) {
	for (long long own = threadIdx.x; own < L.nown; own += blockDim.x) {
		int x = (int)(own % L.nx);
		long long rest = own / L.nx;
		int y = (int)(rest % L.ny);
		int lz = (int)(rest / L.ny) + 1;
		if (((x + y + (lz - 1 + L.zpar) + colour) & 1) != 0) { continue; }
		long long c = own + L.sxy;
		float d = L.diag[c];
		if (d > 0.f) {
			L.x[c] = (L.b[c] + lv_offdiag_sum(L, L.x, c)) / d;
		}
	}
	__syncthreads();
}
__device__ __forceinline__ void tail_restrict(const LevelDev &F, const LevelDev &C) {
	for (long long own = threadIdx.x; own < C.nown; own += blockDim.x) {
		int X_ = (int)(own % C.nx);
		long long rest = own / C.nx;
		int Y_ = (int)(rest % C.ny);
		int LZ = (int)(rest / C.ny) + 1;
		float acc = 0.f;
		for (int k = 0; k < 8; ++k) {
			int x = 2 * X_ + (k & 1), y = 2 * Y_ + ((k >> 1) & 1), lz = 2 * (LZ - 1) + ((k >> 2) & 1) + 1;
			if (x >= F.nx || y >= F.ny || lz > F.nzl) { continue; }
			long long c = x + (long long)F.nx * (y + (long long)F.ny * lz);
			float d = F.diag[c];
			if (d <= 0.f) { continue; }
			acc += F.b[c] - (d * F.x[c] - lv_offdiag_sum(F, F.x, c));
		}
		C.b[own + C.sxy] = acc;
		C.x[own + C.sxy] = 0.f;
	}
	__syncthreads();
}
__device__ __forceinline__ void tail_prolong(const LevelDev &F, const LevelDev &C) {
	for (long long own = threadIdx.x; own < F.nown; own += blockDim.x) {
		int x = (int)(own % F.nx);
		long long rest = own / F.nx;
		int y = (int)(rest % F.ny);
		int lz = (int)(rest / F.ny) + 1;
		long long c = own + F.sxy;
		if (F.diag[c] <= 0.f) { continue; }
		long long cc = (x >> 1) + (long long)C.nx * ((y >> 1) + (long long)C.ny * (((lz - 1) >> 1) + 1));
		F.x[c] += MG_OMEGA * C.x[cc];
	}
	__syncthreads();
}

__global__ void __launch_bounds__(1024) k_mg_tail(TailLevels T, const PcgScalars *scal) {
	if (scal->done) { return; }
	const int last = T.n - 1;
	for (int l = 0; l < last; ++l) { // down: x of the entry level was zeroed by the restriction that filled its b
		for (int s = 0; s < MG_PRE; ++s) {
			tail_half_sweep(T.L[l], 0);
			tail_half_sweep(T.L[l], 1);
		}
		tail_restrict(T.L[l], T.L[l + 1]);
	}
	for (int s = 0; s < MG_COARSE_SWEEPS; ++s) { // coarsest: red, black, black, red
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

// ---- transfer operators -----------------------------------------------------------------------------------------
// coarse b = P^T (b - A x) of the finer level, coarse x = 0.  One thread per coarse cell.
__global__ void __launch_bounds__(128) k_mg_restrict_l0(GridDesc G, const uint8_t *__restrict__ flags,
	const float *__restrict__ b, const float *__restrict__ X, LevelDev C, const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= C.nown) { return; }
	int X_ = (int)(own % C.nx);
	long long rest = own / C.nx;
	int Y_ = (int)(rest % C.ny);
	int LZ = (int)(rest / C.ny) + 1;
	float acc = 0.f;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		int x = 2 * X_ + (k & 1), y = 2 * Y_ + ((k >> 1) & 1), lz = 2 * (LZ - 1) + ((k >> 2) & 1) + 1;
		if (x >= G.nx || y >= G.ny || lz > G.nzl) { continue; }
		long long c = x + (long long)G.nx * (y + (long long)G.ny * lz);
		unsigned f = flags[c];
		if (!(f & FL_L)) { continue; }
		acc += b[c] - ((float)FL_N(f) * X[c] - l0_offdiag_sum(G, f, X, c, x, y));
	}
	long long cc = own + C.sxy;
	C.b[cc] = acc;
	C.x[cc] = 0.f;
}
__global__ void __launch_bounds__(128) k_mg_restrict(LevelDev F, LevelDev C, const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= C.nown) { return; }
	int X_ = (int)(own % C.nx);
	long long rest = own / C.nx;
	int Y_ = (int)(rest % C.ny);
	int LZ = (int)(rest / C.ny) + 1;
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
	long long cc = own + C.sxy;
	C.b[cc] = acc;
	C.x[cc] = 0.f;
}
// x_fine += omega * P x_coarse
__global__ void k_mg_prolong_l0(GridDesc G, const uint8_t *__restrict__ flags, float *__restrict__ X, LevelDev C,
	const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	int x = (int)(own % G.nx);
	long long rest = own / G.nx;
	int y = (int)(rest % G.ny);
	int lz = (int)(rest / G.ny) + 1;
	long long c = own + G.sxy;
	if (!(flags[c] & FL_L)) { return; }
	long long cc = (x >> 1) + (long long)C.nx * ((y >> 1) + (long long)C.ny * (((lz - 1) >> 1) + 1));
	X[c] += MG_OMEGA * C.x[cc];
}
__global__ void k_mg_prolong(LevelDev F, LevelDev C, const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= F.nown) { return; }
	int x = (int)(own % F.nx);
	long long rest = own / F.nx;
	int y = (int)(rest % F.ny);
	int lz = (int)(rest / F.ny) + 1;
	long long c = own + F.sxy;
	if (F.diag[c] <= 0.f) { return; }
	long long cc = (x >> 1) + (long long)C.nx * ((y >> 1) + (long long)C.ny * (((lz - 1) >> 1) + 1));
	F.x[c] += MG_OMEGA * C.x[cc];
}

// ---- Galerkin coarse operators ----------------------------------------------------------------------------------
struct LevelOut {
	float *diag, *cx, *cy, *cz;
};
__global__ void __launch_bounds__(128) k_mg_build_l1(GridDesc G, const uint8_t *__restrict__ flags, LevelDev C,
	LevelOut O) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= C.nown) { return; }
	int X_ = (int)(own % C.nx);
	long long rest = own / C.nx;
	int Y_ = (int)(rest % C.ny);
	int LZ = (int)(rest / C.ny) + 1;
	float diag = 0.f, cxp = 0.f, cyp = 0.f, czp = 0.f;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
		int x = 2 * X_ + dx, y = 2 * Y_ + dy, lz = 2 * (LZ - 1) + dz + 1;
		if (x >= G.nx || y >= G.ny || lz > G.nzl) { continue; }
		long long c = x + (long long)G.nx * (y + (long long)G.ny * lz);
		unsigned f = flags[c];
		if (!(f & FL_L)) { continue; }
		diag += (float)FL_N(f);
		// coupling(c, c + e) exists iff both are unknowns and the + neighbour's type is fluid
		if ((f & FL_XP) && (flags[c + 1] & FL_L)) { if (dx == 0) { diag -= 2.f; } else { cxp += 1.f; } }
		if ((f & FL_YP) && (flags[c + G.nx] & FL_L)) { if (dy == 0) { diag -= 2.f; } else { cyp += 1.f; } }
		if ((f & FL_ZP) && (flags[c + G.sxy] & FL_L)) {
			if (dz == 0 && lz + 1 <= G.nzl) { diag -= 2.f; } else { czp += 1.f; }
		}
	}
	long long cc = own + C.sxy;
	O.diag[cc] = diag;
	O.cx[cc] = cxp;
	O.cy[cc] = cyp;
	O.cz[cc] = czp;
}
__global__ void __launch_bounds__(128) k_mg_build(LevelDev F, LevelDev C, LevelOut O) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= C.nown) { return; }
	int X_ = (int)(own % C.nx);
	long long rest = own / C.nx;
	int Y_ = (int)(rest % C.ny);
	int LZ = (int)(rest / C.ny) + 1;
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
	long long cc = own + C.sxy;
	O.diag[cc] = diag;
	O.cx[cc] = cxp;
	O.cy[cc] = cyp;
	O.cz[cc] = czp;
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int mg_alloc(lfk_ctx *c) {
	if (!c->mg.empty()) { return 0; }
	const GridDesc &G = c->g;
	int nx = G.nx, ny = G.ny, nzl = G.nzl, z0 = G.z0, nz = G.nz;
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
		// slabs must stay aligned to the aggregates; stop coarsening when they would not (multi-GPU)
		bool aligned = c->nranks == 1 || (z0 % 2 == 0 && nzl % 2 == 0 && nz % 2 == 0 && nzl >= 2);
		if (!aligned || (nx <= 2 && ny <= 2 && nzl <= 2)) { break; }
		nx = (nx + 1) / 2; ny = (ny + 1) / 2; nzl = (nzl + 1) / 2; z0 /= 2; nz = (nz + 1) / 2;
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
	c->mg.clear();
	c->mg_z0.clear();
	return 0;
}

int lfkm_setup(lfk_ctx *c, double a_scale) {
	(void)a_scale;
	LFK_TRY(mg_alloc(c));
	const GridDesc &G = c->g;
	for (size_t l = 1; l < c->mg.size(); ++l) {
		MgLevel &C = c->mg[l];
		LevelDev Cd = level_dev(C, c->mg_z0[l]);
		LevelOut O{ C.diag, C.cx, C.cy, C.cz };
		unsigned nb = lfk_blocks(Cd.nown, 128);
		if (l == 1) {
			LFK_LAUNCH(c, k_mg_build_l1, nb, 128, 0, G, c->flags, Cd, O);
		} else {
			LFK_LAUNCH(c, k_mg_build, nb, 128, 0, level_dev(c->mg[l - 1], c->mg_z0[l - 1]), Cd, O);
		}
		if (c->nranks > 1) { // -z couplings and the neighbour's diagonal live in the ghost layers
			LFK_TRY(lfkx_halo_f32(c, C.cz, C.nx, C.ny, C.nzl));
			LFK_TRY(lfkx_halo_f32(c, C.diag, C.nx, C.ny, C.nzl));
		}
	}
	c->mg_valid = true;
	return 0;
}

static int smooth(lfk_ctx *c, size_t l, int first_colour, int sweeps, bool skip_first = false) {
	const GridDesc &G = c->g;
	MgLevel &L = c->mg[l];
	LevelDev Ld = level_dev(L, c->mg_z0[l]);
	long long half = (long long)((L.nx + 1) / 2) * L.ny * L.nzl;
	unsigned nb = lfk_blocks(half, 256);
	for (int s = 0; s < sweeps; ++s) {
		for (int h = 0; h < 2; ++h) {
			int colour = first_colour ^ h;
			if (skip_first && s == 0 && h == 0) { continue; } // already done by the kernel that produced b0 / x0
			if (c->nranks > 1) { LFK_TRY(lfkx_halo_f32(c, L.x, L.nx, L.ny, L.nzl)); }
			if (l == 0) {
				LFK_LAUNCH(c, k_mg_rbgs_l0, nb, 256, 0, G, c->flags, L.b, L.x, colour, c->d_scal);
			} else {
				LFK_LAUNCH(c, k_mg_rbgs, nb, 256, 0, Ld, colour, c->d_scal);
			}
		}
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
		for (size_t k = l; k <= last; ++k) {
			T.L[T.n++] = level_dev(c->mg[k], c->mg_z0[k]);
		}
		LFK_LAUNCH(c, k_mg_tail, 1, 1024, 0, T, c->d_scal);
		return 0;
	}
	if (l == last) { // coarsest level reached outside the tail (multi-GPU alignment limit, or a tiny fine grid)
		LFK_TRY(smooth(c, l, 0, MG_COARSE_SWEEPS));
		LFK_TRY(smooth(c, l, 1, MG_COARSE_SWEEPS));
		return 0;
	}
	LFK_TRY(smooth(c, l, 0, MG_PRE, l == 0)); // red, black (level 0: the first red half-sweep is pre-applied)
	MgLevel &C = c->mg[l + 1];
	LevelDev Cd = level_dev(C, c->mg_z0[l + 1]);
	if (c->nranks > 1) { LFK_TRY(lfkx_halo_f32(c, L.x, L.nx, L.ny, L.nzl)); }
	if (l == 0) {
		LFK_LAUNCH(c, k_mg_restrict_l0, lfk_blocks(Cd.nown, 128), 128, 0, G, c->flags, L.b, L.x, Cd, c->d_scal);
	} else {
		LFK_LAUNCH(c, k_mg_restrict, lfk_blocks(Cd.nown, 128), 128, 0, Ld, Cd, c->d_scal);
	}
	LFK_TRY(vcycle(c, l + 1));
	if (l == 0) {
		LFK_LAUNCH(c, k_mg_prolong_l0, lfk_blocks(G.nown, 256), 256, 0, G, c->flags, L.x, Cd, c->d_scal);
	} else {
		LFK_LAUNCH(c, k_mg_prolong, lfk_blocks(Ld.nown, 256), 256, 0, Ld, Cd, c->d_scal);
	}
	LFK_TRY(smooth(c, l, 1, MG_POST)); // black, red
	return 0;
}

// level-0 buffers, filled by the PCG kernels that produce r (see MgPreload in pressure.cu)
int lfkm_level0(lfk_ctx *c, float **b0, float **x0) {
	LFK_TRY(mg_alloc(c));
	*b0 = c->mg[0].b;
	*x0 = c->mg[0].x;
	return 0;
}

// z = x0 / a_scale (fp64) fused with sigma_new = z.r and its finaliser
__global__ void __launch_bounds__(RED_THREADS) k_mg_store_dot(GridDesc G, const float *__restrict__ x0,
	const uint8_t *__restrict__ flags, const double *__restrict__ r, double *__restrict__ z, double inv_a_scale,
	PcgScalars *scal, double *partials, unsigned *ticket, int finalize, int first) {
	if (scal->done) { return; }
	double acc = 0.0;
	for (long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x; own < G.nown;
		own += (long long)gridDim.x * blockDim.x) {
		long long c = own + G.sxy;
		double zv = (flags[c] & FL_L) ? (double)x0[c] * inv_a_scale : 0.0;
		z[c] = zv;
		acc += zv * r[c];
	}
	acc = block_sum(acc);
	if (threadIdx.x == 0) { partials[blockIdx.x] = acc; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 0);
		if (threadIdx.x == 0) {
			scal->sigma_new = tot;
			if (finalize) { pcg_finalize(scal, first ? FIN_BETA_FIRST : FIN_BETA, 0.0); }
		}
	}
}

// z = M^-1 r, sigma_new = z.r.  b0 and the red half of x0 were written by k_pcg_init / k_update_pr.
int lfkm_apply_preloaded(lfk_ctx *c, double a_scale, unsigned nb, int fin, int first) {
	const GridDesc &G = c->g;
	MgLevel &L0 = c->mg[0];
	LFK_TRY(vcycle(c, 0));
	LFK_LAUNCH(c, k_mg_store_dot, nb, RED_THREADS, 0, G, L0.x, c->flags, c->r, c->z, 1.0 / a_scale, c->d_scal,
		c->partials, c->ticket, fin, first);
	return 0;
}
