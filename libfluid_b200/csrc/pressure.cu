// MAC-grid pressure projection on the dense (slab-local) grid: matrix flags + right-hand side, matrix-free
// 7-point PCG with device-resident scalars, pressure-gradient update, velocity extrapolation.
// Reference: fluid::pressure_solver (src/pressure_solver.cpp:14-371), simulation::_extrapolate_velocities
// (src/simulation.cpp:685-754).
//
// Unknowns live on the dense cell grid (vectors are indexed by local raw cell index and are ZERO on every cell
// that is not in the reference's fluid-cell list), so the stencil needs no index map: the 8-byte-per-cell
// ordinal grid of the reference (pressure_solver.h:47-48) disappears and neighbour access is coalesced.
#include "lfk_internal.cuh"

#include <cmath>
#include <cstring>

// (the flag byte FL_* and the reduction helpers live in lfk_internal.cuh)

__device__ __forceinline__ void cell_xyz(const GridDesc &G, long long own, int &x, int &y, int &lz) {
	x = (int)(own % G.nx);
	long long rest = own / G.nx;
	y = (int)(rest % G.ny);
	lz = (int)(rest / G.ny) + 1;
}

// ---- S1-S3: flags + b (src/pressure_solver.cpp:150-242) ---------------------------------------------------
__global__ void __launch_bounds__(256) k_build_system(GridDesc G, const uint32_t *__restrict__ cnt,
	const uint8_t *__restrict__ typ, const double *__restrict__ u, const double *__restrict__ v,
	const double *__restrict__ w, uint8_t *__restrict__ flags, double *__restrict__ b, double *__restrict__ p,
	double inv_h, double warm_scale) {
	for_own_cells(G, [&](int x, int y, int lz, long long c) {
		if (cnt[c] == 0) {
			p[c] = 0.0;
			flags[c] = 0;
			b[c] = 0.0;
			return;
		}
		// initial guess: 0 like the reference (src/pressure_solver.cpp:24), or -- fused step only -- the previous
		// step's pressure rescaled to this step's dt (p ~ 1 / dt); any finite start converges to the same tolerance
		double p0 = 0.0;
		if (warm_scale != 0.0) {
			p0 = p[c] * warm_scale;
			if (!(fabs(p0) < 1e300)) { p0 = 0.0; }
		}
		p[c] = p0;
		const uint8_t S = LFK_CELL_SOLID, F = LFK_CELL_FLUID;
		// out-of-grid neighbours read as solid (mac_grid::get_cell_and_type); the z ghost layers carry that already
		uint8_t txp = x + 1 < G.nx ? typ[c + 1] : S, txn = x > 0 ? typ[c - 1] : S;
		uint8_t typ_ = y + 1 < G.ny ? typ[c + G.nx] : S, tyn = y > 0 ? typ[c - G.nx] : S;
		uint8_t tzp = typ[c + G.sxy], tzn = typ[c - G.sxy];
		unsigned n = (txp != S) + (typ_ != S) + (tzp != S) + (txn != S) + (tyn != S) + (tzn != S);
		unsigned f = n | FL_L;
		if (typ[c] == F) { f |= FL_SELF; }
		if (txp == F) { f |= FL_XP; }
		if (typ_ == F) { f |= FL_YP; }
		if (tzp == F) { f |= FL_ZP; }
		flags[c] = (uint8_t)f;

		double vx = u[c], vy = v[c], vz = w[c];
		double value = -(vx + vy + vz);
		int z = lz - 1 + G.z0;
		if (x > 0) {
			double q = u[c - 1];
			value += q;
			if (txn == S) { value -= q; }
		}
		if (y > 0) {
			double q = v[c - G.nx];
			value += q;
			if (tyn == S) { value -= q; }
		}
		if (z > 0) {
			double q = w[c - G.sxy];
			value += q;
			if (tzn == S) { value -= q; }
		}
		if (txp == S) { value += vx; }
		if (typ_ == S) { value += vy; }
		if (tzp == S) { value += vz; }
		b[c] = inv_h * value;
	});
}

// ---- S6: out = a_scale * A v (src/pressure_solver.cpp:334-362), same subtraction order as the reference ------
// The six neighbour loads are unconditional (ghost layers / contiguous rows keep every address inside the allocation)
// and masked afterwards, so they are all in flight together with the flag byte; subtracting 0.0 is exact, so the
// result is bit-identical to the reference's conditional form.
struct Stencil7 {
	double c, xm, xp, ym, yp, zm, zp;
	unsigned f;
};
__device__ __forceinline__ Stencil7 stencil_load(const GridDesc &G, const uint8_t *__restrict__ flags,
	const double *__restrict__ s, long long c) {
	Stencil7 v;
	v.f = flags[c];
	v.c = s[c];
	v.xm = s[c - 1];
	v.xp = s[c + 1];
	v.ym = s[c - G.nx];
	v.yp = s[c + G.nx];
	v.zm = s[c - G.sxy];
	v.zp = s[c + G.sxy];
	return v;
}
__device__ __forceinline__ double stencil_apply(const Stencil7 &v, int x, int y, double a_scale) {
	const unsigned f = v.f;
	const bool self = f & FL_SELF; // coupling to -neighbours is the neighbour's "fluid_pos" flag == type(self) == fluid
	double value = (double)FL_N(f) * v.c;
	value -= (self && x > 0) ? v.xm : 0.0;
	value -= (self && y > 0) ? v.ym : 0.0;
	value -= self ? v.zm : 0.0;
	value -= (f & FL_XP) ? v.xp : 0.0;
	value -= (f & FL_YP) ? v.yp : 0.0;
	value -= (f & FL_ZP) ? v.zp : 0.0;
	return a_scale * value;
}

// cells in flight per thread (rows_pipelined) of the two fp64 stencil / update passes
#ifndef SPMV_U
#define SPMV_U 4
#endif
#ifndef UPD_U
#define UPD_U 4
#endif

__global__ void k_finalize(PcgScalars *scal, int which) {
	if (which != FIN_BB && scal->done) { return; }
	pcg_finalize(scal, which);
}

__global__ void __launch_bounds__(RED_THREADS) k_spmv_dot(GridDesc G, const uint8_t *__restrict__ flags,
	const double *__restrict__ s, double *__restrict__ z, PcgScalars *scal,
	double *partials, unsigned *ticket, int finalize) {
	if (scal->done) { return; }
	const double a_scale = scal->a_scale;
	double acc = 0.0;
	rows_pipelined<SPMV_U, Stencil7>(G.nx, G.ny, G.nzl, 0, -1,
		[&](int, int, int, long long c) { return stencil_load(G, flags, s, c); },
		[&](int x, int y, int, long long c, const Stencil7 &v) {
			const double out = (v.f & FL_L) ? stencil_apply(v, x, y, a_scale) : 0.0;
			acc += out * v.c; // out == 0 on cells that are not unknowns
			z[c] = out;
		});
	acc = block_sum(acc);
	if (threadIdx.x == 0) { partials[blockIdx.x] = acc; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 0);
		if (threadIdx.x == 0) {
			scal->zs = tot;
			if (finalize) { pcg_finalize(scal, FIN_ALPHA); }
		}
	}
}

__global__ void __launch_bounds__(256) k_spmv_plain(GridDesc G, const uint8_t *__restrict__ flags,
	const double *__restrict__ s, double *__restrict__ z, double a_scale) {
	for_own_cells(G, [&](int x, int y, int lz, long long c) {
		const Stencil7 v = stencil_load(G, flags, s, c);
		z[c] = (v.f & FL_L) ? stencil_apply(v, x, y, a_scale) : 0.0;
	});
}

// Multigrid input fused into the kernels that produce r: b0 = r / a_scale in fp32, and x0 = the result of the first
// (red) Gauss-Seidel half-sweep from a zero initial guess, i.e. b0 / n on red cells and 0 on black ones.  Saves the
// load kernel, one half-sweep and a re-read of r per PCG iteration.
struct MgPreload {
	float *b0, *x0; // level-0 rhs / solution, or NULL when the preconditioner is not multigrid
	const uint8_t *flags;
};
__device__ __forceinline__ void mg_preload(const GridDesc &G, const MgPreload &M, int x, int y, int lz, long long c,
	double rv, unsigned f, double inv_a_scale) {
	if (M.b0 == nullptr) { return; }
	float bv = (float)(rv * inv_a_scale);
	M.b0[c] = bv;
	bool red = ((x + y + (lz - 1 + G.z0)) & 1) == 0;
	M.x0[c] = (red && (f & FL_L) && FL_N(f) > 0) ? bv / (float)FL_N(f) : 0.f;
}

// r = b, sum b^2
__global__ void __launch_bounds__(RED_THREADS) k_pcg_init(GridDesc G, const double *__restrict__ b,
	double *__restrict__ r, PcgScalars *scal, double *partials, unsigned *ticket, int finalize, MgPreload M,
	double a_scale, double tolerance) {
	double acc = 0.0;
	const double inv_a_scale = 1.0 / a_scale;
	for_own_cells(G, [&](int x, int y, int lz, long long c) {
		double v = b[c];
		r[c] = v;
		acc += v * v;
		mg_preload(G, M, x, y, lz, c, v, M.flags[c], inv_a_scale);
	});
	acc = block_sum(acc);
	if (threadIdx.x == 0) { partials[blockIdx.x] = acc; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 0);
		if (threadIdx.x == 0) {
			scal->bb = tot;
			scal->sigma = 0.0;
			scal->a_scale = a_scale; // the iteration's kernels read these from here
			scal->inv_a_scale = 1.0 / a_scale;
			scal->tolerance = tolerance;
			if (finalize) { pcg_finalize(scal, FIN_BB); }
		}
	}
}

// warm start: r = b - A p, sum b^2 (the early-out of the reference is decided on b, as there)
__global__ void __launch_bounds__(RED_THREADS) k_pcg_init_warm(GridDesc G, const uint8_t *__restrict__ flags,
	const double *__restrict__ b, const double *__restrict__ p, double *__restrict__ r, double a_scale,
	PcgScalars *scal, double *partials, unsigned *ticket, int finalize, MgPreload M, double tolerance) {
	double acc = 0.0;
	const double inv_a_scale = 1.0 / a_scale;
	for_own_cells(G, [&](int x, int y, int lz, long long c) {
		const Stencil7 v = stencil_load(G, flags, p, c);
		const double bv = b[c];
		const double rv = (v.f & FL_L) ? bv - stencil_apply(v, x, y, a_scale) : bv; // b == 0 off the unknowns
		r[c] = rv;
		acc += bv * bv;
		mg_preload(G, M, x, y, lz, c, rv, v.f, inv_a_scale);
	});
	acc = block_sum(acc);
	if (threadIdx.x == 0) { partials[blockIdx.x] = acc; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 0);
		if (threadIdx.x == 0) {
			scal->bb = tot;
			scal->sigma = 0.0;
			scal->a_scale = a_scale; // the iteration's kernels read these from here
			scal->inv_a_scale = 1.0 / a_scale;
			scal->tolerance = tolerance;
			if (finalize) { pcg_finalize(scal, FIN_BB); }
		}
	}
}

// Jacobi: z = r / (a_scale * n)
__global__ void k_precond_jacobi(GridDesc G, const uint8_t *__restrict__ flags, const double *__restrict__ r,
	double *__restrict__ z, const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long c = own + G.sxy;
	unsigned f = flags[c];
	double out = 0.0;
	if ((f & FL_L) && FL_N(f) > 0) {
		out = r[c] / (scal->a_scale * (double)FL_N(f));
	}
	z[c] = out;
}

// sigma_new = z.r
__global__ void __launch_bounds__(RED_THREADS) k_dot_zr(GridDesc G, const double *__restrict__ z,
	const double *__restrict__ r, PcgScalars *scal, double *partials, unsigned *ticket, int finalize, int first) {
	if (scal->done) { return; }
	double acc = 0.0;
	for_own_cells(G, [&](int, int, int, long long c) { acc += z[c] * r[c]; });
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

// p += alpha s ; r -= alpha z ; residual = max |r|
__global__ void __launch_bounds__(RED_THREADS) k_update_pr(GridDesc G, double *__restrict__ p,
	double *__restrict__ r, const double *__restrict__ s, const double *__restrict__ z, PcgScalars *scal,
	double *partials, unsigned *ticket, int finalize, MgPreload M) {
	if (scal->done) { return; }
	const double alpha = scal->alpha, inv_a_scale = scal->inv_a_scale;
	double m = 0.0;
	struct PR { double p, s, r, z; unsigned f; };
	rows_pipelined<UPD_U, PR>(G.nx, G.ny, G.nzl, 0, -1,
		[&](int, int, int, long long c) {
			PR v;
			v.p = p[c];
			v.s = s[c];
			v.r = r[c];
			v.z = z[c];
			v.f = M.flags[c];
			return v;
		},
		[&](int x, int y, int lz, long long c, const PR &v) {
			const double rv = v.r + (-alpha) * v.z;
			p[c] = v.p + alpha * v.s;
			r[c] = rv;
			m = fmax(m, fabs(rv));
			mg_preload(G, M, x, y, lz, c, rv, v.f, inv_a_scale);
		});
	m = block_max(m);
	if (threadIdx.x == 0) { partials[blockIdx.x] = m; }
	if (lfk_last_block(ticket)) {
		double tot = finish_partials(partials, gridDim.x, 1);
		if (threadIdx.x == 0) {
			scal->resmax = tot;
			if (finalize) { pcg_finalize(scal, FIN_RESID); }
		}
	}
}

// s = z + beta s  (first iteration: beta == 0 => s = z)
__global__ void k_xpby(GridDesc G, double *__restrict__ s, const double *__restrict__ z, const PcgScalars *scal) {
	if (scal->done) { return; }
	const double beta = scal->beta;
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long c = own + G.sxy;
	s[c] = beta == 0.0 ? z[c] : z[c] + beta * s[c];
}

int lfks_build_system(lfk_ctx *c, double dt) {
	PhaseTimer T(c, LFK_PHASE_SOLVE_SETUP);
	const GridDesc &G = c->g;
	if (c->nranks > 1) {
		LFK_TRY(lfkx_halo_u8(c, c->typ));
		for (int d = 0; d < 3; ++d) { LFK_TRY(lfkx_halo_f64(c, c->vel[d])); }
	}
	LFK_LAUNCH(c, k_build_system, lfk_row_blocks(G, 256, 1u << 20), 256, 0, G, c->cnt, c->typ, c->vel[0], c->vel[1],
		c->vel[2], c->flags, c->b, c->p, 1.0 / G.h, c->warm_scale);
	c->warm_applied = c->warm_scale != 0.0;
	c->warm_scale = 0.0;
	if (c->nranks > 1) {
		LFK_TRY(lfkx_halo_u8(c, c->flags));
	}
	c->system_valid = true;
	c->system_dt = dt;
	c->mg_valid = false;
	return 0;
}

// Grid of the PCG reduction kernels (SpMV + dot, p / r update, final sweep + dot): persistent-style, 8 blocks per SM,
// rows strided over the warps -- measured at 256^3 (profiles/r1c_sweep.json): 1.01 ms / iteration against 1.11 with one
// row per warp (16384 blocks: more block partials for the last block to reduce, a longer tail).
static inline unsigned red_blocks(const lfk_ctx *c) {
	static int sms = 0;
	if (sms == 0) {
		if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || sms <= 0) { sms = 148; }
	}
	const unsigned cap = c->tune.red_blocks > 0 ? (unsigned)c->tune.red_blocks : 8u * (unsigned)sms;
	return lfk_row_blocks(c->g, RED_THREADS, cap < RED_BLOCKS ? cap : RED_BLOCKS);
}

static int allreduce_scalar(lfk_ctx *c, double *field, bool is_max, int which) {
	if (c->nranks > 1) {
		if (lfkx_allreduce_finalize(c, field, is_max, which)) { return 0; } // one kernel over peer memory
		LFK_TRY(is_max ? lfkx_allreduce_max(c, field, 1) : lfkx_allreduce_sum(c, field, 1));
		LFK_LAUNCH(c, k_finalize, 1, 1, 0, c->d_scal, which);
	}
	return 0;
}

int lfks_free_graph(lfk_ctx *c);
int lfkm_setup(lfk_ctx *c, double a_scale);                                    // mg.cu
int lfkm_level0(lfk_ctx *c, float **b0, float **x0);                           // mg.cu
int lfkm_apply_preloaded(lfk_ctx *c, unsigned nb, int fin, int first); // mg.cu: V-cycle + (z, z.r)

// z = M^-1 r and sigma_new = z.r (+ its finaliser on one GPU)
static int precondition_and_dot(lfk_ctx *c, unsigned nb, int fin, int first) {
	const GridDesc &G = c->g;
	if (c->prm.preconditioner == LFK_PRECOND_MULTIGRID) {
		return lfkm_apply_preloaded(c, nb, fin, first);
	}
	LFK_LAUNCH(c, k_precond_jacobi, lfk_blocks(G.nown, 256), 256, 0, G, c->flags, c->r, c->z, c->d_scal);
	LFK_LAUNCH(c, k_dot_zr, nb, RED_THREADS, 0, G, c->z, c->r, c->d_scal, c->partials, c->ticket, fin, first);
	return 0;
}

// pressure_solver::solve (src/pressure_solver.cpp:19-71)
int lfks_solve(lfk_ctx *c, double dt, double *residual, uint64_t *iters, bool warm) {
	const GridDesc &G = c->g;
	// warm start (fused step only): the previous solve's pressure, rescaled by dt_prev / dt, is the initial guess
	c->warm_applied = false;
	if (warm && c->tune.warm_start && c->last_solve_dt > 0.0 && dt > 0.0 && c->last_solve_ok) {
		c->warm_scale = c->last_solve_dt / dt;
		c->system_valid = false; // the initial guess is written by the kernel that builds the system
	}
	if (!c->system_valid || c->system_dt != dt) {
		LFK_TRY(lfks_build_system(c, dt));
	} else if (c->pressure_valid) { // the system is reused and p holds the previous solution: start from 0 again
		LFK_CUDA(c, cudaMemsetAsync(c->p, 0, (size_t)G.ncl * sizeof(double), c->stream));
	}
	c->warm_scale = 0.0;
	const bool warmed = c->warm_applied;
	PhaseTimer T(c, LFK_PHASE_PCG);
	const double a_scale = dt / (c->prm.density * G.h * G.h);
	const double tol = c->prm.tolerance;
	const int fin = c->nranks == 1 ? 1 : 0;
	const unsigned nb = red_blocks(c), eb = lfk_blocks(G.nown, 256);
	if (c->prm.preconditioner == LFK_PRECOND_MULTIGRID && !c->mg_valid) {
		LFK_TRY(lfkm_setup(c, a_scale));
	}
	MgPreload M{ nullptr, nullptr, c->flags };
	if (c->prm.preconditioner == LFK_PRECOND_MULTIGRID) {
		LFK_TRY(lfkm_level0(c, &M.b0, &M.x0));
	}
	if (warmed) {
		if (c->nranks > 1) { LFK_TRY(lfkx_halo_f64(c, c->p)); }
		LFK_LAUNCH(c, k_pcg_init_warm, nb, RED_THREADS, 0, G, c->flags, c->b, c->p, c->r, a_scale, c->d_scal, c->partials,
			c->ticket, fin, M, tol);
	} else {
		LFK_LAUNCH(c, k_pcg_init, nb, RED_THREADS, 0, G, c->b, c->r, c->d_scal, c->partials, c->ticket, fin, M, a_scale, tol);
	}
	LFK_TRY(allreduce_scalar(c, &c->d_scal->bb, false, FIN_BB));
	LFK_TRY(precondition_and_dot(c, nb, fin, 1));
	LFK_TRY(allreduce_scalar(c, &c->d_scal->sigma_new, false, FIN_BETA_FIRST));
	LFK_LAUNCH(c, k_xpby, eb, 256, 0, G, c->s, c->z, c->d_scal);

	// one iteration (src/pressure_solver.cpp:44-68); with_direction: everything after the residual test
	auto iteration = [&](bool with_direction) -> int {
		if (c->nranks > 1) { LFK_TRY(lfkx_halo_f64(c, c->s)); }
		LFK_LAUNCH(c, k_spmv_dot, nb, RED_THREADS, 0, G, c->flags, c->s, c->z, c->d_scal, c->partials, c->ticket, fin);
		LFK_TRY(allreduce_scalar(c, &c->d_scal->zs, false, FIN_ALPHA));
		LFK_LAUNCH(c, k_update_pr, nb, RED_THREADS, 0, G, c->p, c->r, c->s, c->z, c->d_scal, c->partials, c->ticket, fin, M);
		LFK_TRY(allreduce_scalar(c, &c->d_scal->resmax, true, FIN_RESID));
		if (with_direction) { // the reference leaves the loop after max_iterations updates of p, r
			LFK_TRY(precondition_and_dot(c, nb, fin, 0));
			LFK_TRY(allreduce_scalar(c, &c->d_scal->sigma_new, false, FIN_BETA));
			LFK_LAUNCH(c, k_xpby, eb, 256, 0, G, c->s, c->z, c->d_scal);
		}
		return 0;
	};
	// The iteration's kernels take every per-solve quantity from device memory (PcgScalars), so ONE captured CUDA graph
	// serves every iteration of every solve of this context; it is re-captured only when the launch configuration
	// changes.  ~40 launches per iteration collapse into one graph launch (launch gaps were ~17 % of the iteration).
	// (multi-GPU: the peer-memory halo / all-reduce kernels and the NCCL calls are captured like any other node)
	const bool use_graph = c->tune.graph != 0;
	const int graph_key = (((int)nb * 2 + c->prm.preconditioner) * 2 + (c->tune.p2p ? 1 : 0)) * 4 + (c->tune.mg_agg ? 2 : 0) + 1 + 4099 * c->tune.ll_kb;
	if (use_graph && (c->pcg_graph == nullptr || c->pcg_graph_key != graph_key)) {
		LFK_TRY(lfks_free_graph(c));
		cudaGraph_t graph = nullptr;
		const uint64_t before = c->stats.kernel_launches;
		LFK_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
		const int rc_it = iteration(true);
		const cudaError_t e_end = cudaStreamEndCapture(c->stream, &graph);
		if (rc_it != 0 || e_end != cudaSuccess || graph == nullptr) {
			if (graph) { cudaGraphDestroy(graph); }
			cudaGetLastError();
			return lfk_fail(c, rc_it != 0 ? rc_it : -(int)e_end, "capturing the PCG iteration graph failed", __FILE__, __LINE__);
		}
		c->pcg_graph_launches = (unsigned)(c->stats.kernel_launches - before);
		c->stats.kernel_launches = before; // captured, not launched
		cudaGraphExec_t exec = nullptr;
		const cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
		cudaGraphDestroy(graph);
		LFK_CUDA(c, e_inst);
		c->pcg_graph = exec;
		c->pcg_graph_key = graph_key;
	}

	// The loop runs entirely from device-resident scalars; the host only polls the `done` flag now and then.
	// Kernels issued after convergence return immediately, so the iteration count stays exact.
	const int max_it = c->prm.max_iterations;
	// every rank must issue the same number of iterations (the NCCL calls inside them pair up across ranks), so the
	// polling interval is derived from the WHOLE grid, never from the rank's own slab (slabs may differ by a layer)
	const long long per_rank = G.sxy * (long long)G.nz / c->nranks;
	int poll = per_rank >= (1ll << 22) ? 2 : (per_rank >= (1ll << 18) ? 8 : 16);
	int issued = 0;
	bool done = max_it <= 0;
	// first burst: one short of what the previous solve of this context needed (consecutive steps need about the same
	// number of iterations), so that at most a couple of no-op iterations are issued past convergence
	int first_burst = (int)c->last_iters - 1;
	while (!done) {
		int want = issued == 0 && first_burst > poll ? first_burst : poll;
		int burst = max_it - issued < want ? max_it - issued : want;
		for (int k = 0; k < burst; ++k) {
			const bool with_direction = issued + k + 1 < max_it;
			if (use_graph && with_direction) {
				LFK_CUDA(c, cudaGraphLaunch((cudaGraphExec_t)c->pcg_graph, c->stream));
				c->stats.kernel_launches += c->pcg_graph_launches;
			} else {
				LFK_TRY(iteration(with_direction));
			}
		}
		issued += burst;
		LFK_TRY(lfk_readback(c, c->h_scal, c->d_scal, sizeof(PcgScalars)));
		LFK_CUDA(c, cudaStreamSynchronize(c->stream));
		done = c->h_scal->done != 0 || issued >= max_it;
		if (poll < 8 && first_burst <= 2) { poll *= 2; }
	}
	if (c->nranks > 1) { LFK_TRY(lfkx_check(c)); }
	if (warmed && c->h_scal->iters == 0 && c->h_scal->bb < 1e-6) { // early-out of the reference: p = 0 (:29-35)
		LFK_CUDA(c, cudaMemsetAsync(c->p, 0, (size_t)G.ncl * sizeof(double), c->stream));
	}
	c->last_iters = c->h_scal->iters;
	c->last_solve_dt = dt;
	c->last_solve_ok = c->h_scal->done != 0 && c->h_scal->resmax < 1e300;
	c->stats.pcg_iterations = c->h_scal->iters;
	c->stats.pcg_residual = c->h_scal->resmax;
	if (residual) { *residual = c->h_scal->resmax; }
	if (iters) { *iters = c->h_scal->iters; }
	c->pressure_valid = true;
	return 0;
}

// out = A v on dense vectors (parity hook for _apply_a)
int lfks_free_graph(lfk_ctx *c) {
	if (c->pcg_graph) {
		cudaGraphExecDestroy((cudaGraphExec_t)c->pcg_graph);
		c->pcg_graph = nullptr;
	}
	return 0;
}

int lfks_apply_a(lfk_ctx *c, double dt, const double *d_v_dense, double *d_out_dense) {
	const GridDesc &G = c->g;
	if (!c->system_valid || c->system_dt != dt) {
		LFK_TRY(lfks_build_system(c, dt));
	}
	const double a_scale = dt / (c->prm.density * G.h * G.h);
	LFK_LAUNCH(c, k_spmv_plain, lfk_row_blocks(G, 256, 1u << 20), 256, 0, G, c->flags, d_v_dense, d_out_dense, a_scale);
	return 0;
}

// ---- S9: pressure-gradient update, face-centric (src/pressure_solver.cpp:73-148) ---------------------------
// Every +face is owned by exactly one cell c; the reference touches it from c (if c is an unknown: "+face" rule)
// and then from the neighbour d = c + e (if d is an unknown: "-face" rule), in that order (c < d in raw order).
__device__ __forceinline__ double face_update(double val, bool Lc, bool Ld, uint8_t tc, uint8_t td, double pc,
	double pd, double coeff) {
	if (Lc) {
		if (td != LFK_CELL_SOLID) {
			double otherp = td == LFK_CELL_FLUID ? pd : 0.0;
			val -= coeff * (otherp - pc);
		} else {
			val = 0.0;
		}
	}
	if (Ld) {
		if (tc == LFK_CELL_AIR) {
			val -= coeff * pd;
		} else if (tc == LFK_CELL_SOLID) {
			val = 0.0;
		}
	}
	return val;
}

__global__ void __launch_bounds__(256) k_apply_pressure(GridDesc G, const uint8_t *__restrict__ flags,
	const uint8_t *__restrict__ typ, const double *__restrict__ p, double *__restrict__ u, double *__restrict__ v,
	double *__restrict__ w, double coeff) {
	for_own_cells(G, [&](int x, int y, int lz, long long c) {
		const bool Lc = flags[c] & FL_L;
		const uint8_t tc = typ[c];
		const double pc = p[c];
		{
			bool in = x + 1 < G.nx;
			bool Ld = in && (flags[c + 1] & FL_L);
			if (Lc || Ld) {
				u[c] = face_update(u[c], Lc, Ld, tc, in ? typ[c + 1] : (uint8_t)LFK_CELL_SOLID, pc, in ? p[c + 1] : 0.0,
					coeff);
			}
		}
		{
			bool in = y + 1 < G.ny;
			bool Ld = in && (flags[c + G.nx] & FL_L);
			if (Lc || Ld) {
				v[c] = face_update(v[c], Lc, Ld, tc, in ? typ[c + G.nx] : (uint8_t)LFK_CELL_SOLID, pc,
					in ? p[c + G.nx] : 0.0, coeff);
			}
		}
		{ // z: the ghost layer is solid / not an unknown at the domain boundary, the neighbour's cells otherwise
			bool Ld = flags[c + G.sxy] & FL_L;
			if (Lc || Ld) {
				w[c] = face_update(w[c], Lc, Ld, tc, typ[c + G.sxy], pc, p[c + G.sxy], coeff);
			}
		}
	});
}

int lfks_apply_pressure(lfk_ctx *c, double dt) {
	PhaseTimer T(c, LFK_PHASE_APPLY_PRESSURE);
	LFK_REQUIRE(c, c->system_valid && c->pressure_valid, LFK_E_STATE,
		"lfk_apply_pressure needs lfk_pressure_solve (or lfk_upload_pressure) first");
	const GridDesc &G = c->g;
	if (c->nranks > 1) { LFK_TRY(lfkx_halo_f64(c, c->p)); }
	double coeff = dt / (c->prm.density * G.h);
	LFK_LAUNCH(c, k_apply_pressure, lfk_row_blocks(G, 256, 1u << 20), 256, 0, G, c->flags, c->typ, c->p, c->vel[0],
		c->vel[1], c->vel[2], coeff);
	c->system_valid = false; // b was built from the pre-projection velocities
	return 0;
}

// ---- E1: velocity extrapolation (src/simulation.cpp:685-754) ----------------------------------------------
__global__ void k_valid_from_counts(long long ncl, const uint32_t *__restrict__ cnt, uint8_t *__restrict__ valid) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ncl) { valid[i] = cnt[i] > 0 ? 1 : 0; }
}

__global__ void k_extrapolate(GridDesc G, const uint8_t *__restrict__ valid, uint8_t *__restrict__ valid_next,
	const uint8_t *__restrict__ typ, double *__restrict__ u, double *__restrict__ v, double *__restrict__ w) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long c = own + G.sxy;
	if (valid[c]) { return; }
	int x, y, lz;
	cell_xyz(G, own, x, y, lz);
	int z = lz - 1 + G.z0;
	int nvalid = 0;
	double s0 = 0.0, s1 = 0.0, s2 = 0.0;
	uint8_t tp0 = LFK_CELL_SOLID, tp1 = LFK_CELL_SOLID, tp2 = LFK_CELL_SOLID;
#define TAKE(nb) do { s0 += u[nb]; s1 += v[nb]; s2 += w[nb]; ++nvalid; } while (0)
	if (x > 0 && valid[c - 1]) { TAKE(c - 1); }
	if (x + 1 < G.nx && valid[c + 1]) { TAKE(c + 1); tp0 = typ[c + 1]; }
	if (y > 0 && valid[c - G.nx]) { TAKE(c - G.nx); }
	if (y + 1 < G.ny && valid[c + G.nx]) { TAKE(c + G.nx); tp1 = typ[c + G.nx]; }
	if (z > 0 && valid[c - G.sxy]) { TAKE(c - G.sxy); }
	if (z + 1 < G.nz && valid[c + G.sxy]) { TAKE(c + G.sxy); tp2 = typ[c + G.sxy]; }
#undef TAKE
	if (nvalid > 0) {
		// only cells WITHOUT particles are written and only cells WITH particles are read: no race in place
		uint8_t t = typ[c];
		double dn = (double)nvalid;
		if (t == tp0) { u[c] = s0 / dn; }
		if (t == tp1) { v[c] = s1 / dn; }
		if (t == tp2) { w[c] = s2 / dn; }
		valid_next[c] = 1;
	}
}

int lfks_extrapolate(lfk_ctx *c) {
	PhaseTimer T(c, LFK_PHASE_EXTRAPOLATE);
	const GridDesc &G = c->g;
	int iters = c->prm.extrapolation_iterations;
	if (iters <= 0) { return 0; }
	LFK_LAUNCH(c, k_valid_from_counts, lfk_blocks(G.ncl, 256), 256, 0, G.ncl, c->cnt, c->valid[0]);
	int cur = 0;
	for (int it = 0; it < iters; ++it) {
		if (c->nranks > 1) {
			LFK_TRY(lfkx_halo_u8(c, c->valid[cur]));
			for (int d = 0; d < 3; ++d) { LFK_TRY(lfkx_halo_f64(c, c->vel[d])); }
		}
		if (it + 1 < iters) {
			LFK_CUDA(c, cudaMemcpyAsync(c->valid[cur ^ 1], c->valid[cur], (size_t)G.ncl, cudaMemcpyDeviceToDevice,
				c->stream));
		}
		LFK_LAUNCH(c, k_extrapolate, lfk_blocks(G.nown, 256), 256, 0, G, c->valid[cur], c->valid[cur ^ 1], c->typ,
			c->vel[0], c->vel[1], c->vel[2]);
		cur ^= 1;
	}
	c->system_valid = false;
	return 0;
}

// ---- fluid-cell ordinals: compaction between the dense grid and the reference's fluid-cell list --------------
int lfks_ensure_ordinal(lfk_ctx *c) {
	if (c->ordinal_valid) { return 0; }
	LFK_TRY(lfkp_exclusive_scan_u32(c, c->cnt, c->ordinal, c->g.ncl, 1));
	c->ordinal_valid = true;
	return 0;
}

__global__ void k_compact_f64(long long ncl, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ ord,
	const double *__restrict__ dense, double *__restrict__ out) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ncl && cnt[i] > 0) { out[ord[i]] = dense[i]; }
}
__global__ void k_compact_flags(long long ncl, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ ord,
	const uint8_t *__restrict__ flags, uint8_t *__restrict__ out) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ncl && cnt[i] > 0) {
		unsigned f = flags[i];
		out[ord[i]] = (uint8_t)(FL_N(f) | (((f >> 5) & 7u) << 3)); // reference cell_data bit order
	}
}
__global__ void k_expand_f64(long long ncl, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ ord,
	const double *__restrict__ in, double *__restrict__ dense) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ncl) { dense[i] = cnt[i] > 0 ? in[ord[i]] : 0.0; }
}
__global__ void k_fluid_cells(long long ncl, long long shift, const uint32_t *__restrict__ cnt,
	const uint32_t *__restrict__ ord, unsigned long long *__restrict__ out) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < ncl && cnt[i] > 0) { out[ord[i]] = (unsigned long long)(i + shift); }
}

int lfks_compact(lfk_ctx *c, const double *dense, double *d_out, const uint8_t *dense_u8, uint8_t *d_out_u8) {
	LFK_TRY(lfks_ensure_ordinal(c));
	unsigned nb = lfk_blocks(c->g.ncl, 256);
	if (dense) { LFK_LAUNCH(c, k_compact_f64, nb, 256, 0, c->g.ncl, c->cnt, c->ordinal, dense, d_out); }
	if (dense_u8) { LFK_LAUNCH(c, k_compact_flags, nb, 256, 0, c->g.ncl, c->cnt, c->ordinal, dense_u8, d_out_u8); }
	return 0;
}
int lfks_expand(lfk_ctx *c, const double *d_compact, double *dense) {
	LFK_TRY(lfks_ensure_ordinal(c));
	LFK_LAUNCH(c, k_expand_f64, lfk_blocks(c->g.ncl, 256), 256, 0, c->g.ncl, c->cnt, c->ordinal, d_compact, dense);
	return 0;
}
int lfks_fluid_cells(lfk_ctx *c, uint64_t *d_out) {
	LFK_TRY(lfks_ensure_ordinal(c));
	long long shift = ((long long)c->g.z0 - 1) * c->g.sxy; // local raw -> whole-grid raw
	LFK_LAUNCH(c, k_fluid_cells, lfk_blocks(c->g.ncl, 256), 256, 0, c->g.ncl, shift, c->cnt, c->ordinal,
		(unsigned long long*)d_out);
	return 0;
}

// ---- config 5: projection-only synthetic system --------------------------------------------------------------
__global__ void k_synthetic_projection(GridDesc G, uint8_t *__restrict__ typ, uint32_t *__restrict__ cnt,
	double *__restrict__ u, double *__restrict__ v, double *__restrict__ w, unsigned long long seed) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	int x, y, lz;
	cell_xyz(G, own, x, y, lz);
	long long c = own + G.sxy;
	long long graw = c + ((long long)G.z0 - 1) * G.sxy;
	bool air = y == G.ny - 1; // free surface: top layer is air => the system is non-singular
	typ[c] = air ? LFK_CELL_AIR : LFK_CELL_FLUID;
	cnt[c] = air ? 0u : 1u;
	unsigned long long r = mix64(seed ^ mix64((unsigned long long)graw + 0x9e3779b97f4a7c15ull));
	double f[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		r = mix64(r + 0x9e3779b97f4a7c15ull);
		f[d] = (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
	}
	u[c] = f[0];
	v[c] = f[1];
	w[c] = f[2];
}

int lfks_synthetic_projection(lfk_ctx *c, uint64_t seed) {
	const GridDesc &G = c->g;
	LFK_CUDA(c, cudaMemsetAsync(c->cnt, 0, (size_t)G.ncl * sizeof(uint32_t), c->stream));
	LFK_LAUNCH(c, k_synthetic_projection, lfk_blocks(G.nown, 256), 256, 0, G, c->typ, c->cnt, c->vel[0], c->vel[1],
		c->vel[2], (unsigned long long)seed);
	c->np = 0;
	c->first = 0;
	c->ntot = 0;
	c->speed2_valid = false;
	c->table_valid = false;
	c->ordinal_valid = false;
	c->system_valid = false;
	c->pressure_valid = false;
	return 0;
}
