// Internal declarations shared by the lfk translation units (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "lfk.h"

// ---------------------------------------------------------------------------------------------------------
// Device-side description of the (slab of the) MAC grid.
//
// Every per-cell device array covers the cells this rank owns PLUS one ghost layer in z on either side:
//   local raw index  lr = x + nx * (y + ny * lz),   lz = z - z0 + 1  in [0, nzl + 2)
// x and y are not padded.  At the outer domain boundary the ghost layers are "outside the grid": type solid,
// velocity 0, no particles -- which is exactly how the reference treats out-of-range neighbours
// (mac_grid::get_cell_and_type, reference src/mac_grid.cpp:26-38).  Between ranks they mirror the neighbour's
// boundary layer.
// ---------------------------------------------------------------------------------------------------------
struct GridDesc {
	int nx, ny, nz;     // global size
	int z0, nzl;        // first owned global z, number of owned layers
	int nlz;            // nzl + 2
	long long sxy;      // nx * ny
	long long ncl;      // nx * ny * nlz  (local cells incl. ghosts)
	long long nown;     // nx * ny * nzl
	double h;           // cell_size
	double inv_h;       // 1 / cell_size
	int hpow2;          // cell_size is a power of two: x / h == x * inv_h bit for bit, so the division is skipped
	double off[3];      // grid_offset
};

enum { PF_PX = 0, PF_PY, PF_PZ, PF_VX, PF_VY, PF_VZ, PF_C0, PF_OX = 15, PF_OY, PF_OZ, PF_COUNT = 18 };

struct ParticleSoA {
	double *f[PF_COUNT];
};

// scalars of the PCG loop, kept on the device so that the iteration needs no host round trip
struct PcgScalars {
	double sigma;      // z.r of the previous iteration
	double zs;         // z.s
	double sigma_new;  // z.r
	double resmax;     // max |r|
	double bb;         // sum b^2
	double alpha, beta;
	double a_scale, inv_a_scale, tolerance; // of the running solve (device resident: the iteration's kernels take no
	                                        // per-solve arguments, so its captured CUDA graph is reused from step to step)
	unsigned long long iters;
	int done;          // 1 once converged / early-out
	int pad;
};

struct MgLevel {
	int nx, ny, nlz;         // local size incl. z ghosts
	int nzl;
	long long sxy, ncl;
	float *diag, *cx, *cy, *cz; // Galerkin 7-point operator (level >= 1); level 0 uses the flags
	float *x, *b, *r;        // solution / rhs / residual scratch (fp32)
};

#define LFK_MAX_DEVICES 64 // per-device one-time set-up flags (cudaFuncSetAttribute is a per-device setting)

// A/B switches and tuning knobs (lfk_set_tuning; the defaults are the production path)
enum { LFK_TUNE_P2G_MARCH = 0, LFK_TUNE_P2G_GATHER = 2 };
#define LFK_REDUCE_SPEED2 8

struct lfk_tuning {
	int p2g = LFK_TUNE_P2G_MARCH; // 2: the plain per-cell gather (the reference's loop literally; also taken for APIC with h < 1)
	int p2g_chunk = 0; // > 0: z planes per block of the marching P2G kernel (8 .. 128; default: ~6 waves of one block per SM)
	int lean_sort = 1; // fused step: 1 the sort permutes positions only and P2G reads velocity / c rows through the
	                   // permutation, 0 the sort permutes the whole payload
	int mg_agg = 1;   // multi-GPU: 1 coarse levels agglomerated onto every rank (r2d: 1.68 against 2.06 ms per iteration on 2 GPUs), 0 distributed
	int mg_agg_cells = 0; // > 0: largest whole-grid level (cells) that is agglomerated (default 600000)
	int mg_coarse = 0; // > 0: symmetric sweeps on the coarsest multigrid level of the single-block tail (default 8)
	int graph = 1;    // 1: the PCG iteration is replayed from a captured CUDA graph (~40 launches), 0: launched one by one
	int ll_kb = 1024; // multi-GPU: layers up to this size (KB, <= 2048) use the flag-in-data halo protocol (4 GPUs,
	                  // 512 x 512 layers: 1.33 ms per PCG iteration at 1024, 1.36 at 256, 1.46 at 64; r2v)
	int p2p = 1;      // multi-GPU: 1 halos through peer memory (CUDA IPC arenas over NVLink), 0 NCCL send / recv
	int warm_start = 1; // fused step: start PCG from the previous step's pressure (0: from p = 0 like the reference)
	int red_blocks = 0; // > 0: cap on the grid of the PCG reduction kernels (default 8 x SM count)
};

struct lfk_ctx {
	int device = 0;
	lfk_tuning tune;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	int nranks = 1, rank = 0;
	void *comm = nullptr; // ncclComm_t
	GridDesc g{};
	lfk_params prm{};
	std::string err;

	// particles: the rank's OWN particles are entries [first, first + np) of the SoA arrays.  first != 0 only between a
	// multi-GPU sort and the next one: the sorted array then reads [ghost copies of the lower neighbour's top layer |
	// own | ghost copies of the upper neighbour's bottom layer], ntot entries in all, addressed by the cell table.
	uint64_t np = 0, cap = 0, first = 0, ntot = 0;
	ParticleSoA P{}, Palt{};
	uint32_t *key = nullptr, *key_alt = nullptr, *slot = nullptr, *perm = nullptr;
	bool old_valid = false;   // false: old_position == position (not materialised)
	bool table_valid = false;
	bool keys_valid = false;
	// lean sort: velocity / c rows are still in the order before the last sort; perm[i] = their index for particle i
	bool v_deferred = false, c_deferred = false;

	// grid (local cells incl. ghosts)
	double *vel[3] = { nullptr, nullptr, nullptr };
	double *vel_old[3] = { nullptr, nullptr, nullptr };
	uint8_t *typ = nullptr;
	uint32_t *cnt = nullptr, *begin = nullptr; // begin has ncl + 1 entries
	double *ctr[3] = { nullptr, nullptr, nullptr }; // cell-centre coordinates per axis (reference: repeated addition)
	uint8_t *valid[2] = { nullptr, nullptr };
	double *wlow[2] = { nullptr, nullptr }; // multi-GPU: w (and FLIP's snapshot) of layer z0 - 2, for G2P (g2p.cu)

	// solver
	uint8_t *flags = nullptr;
	double *b = nullptr, *p = nullptr, *r = nullptr, *z = nullptr, *s = nullptr;
	bool system_valid = false; double system_dt = 0.0;
	bool pressure_valid = false;
	double warm_scale = 0.0;        // != 0: lfks_build_system seeds p with warm_scale * (previous p)
	bool warm_applied = false;
	double last_solve_dt = 0.0; bool last_solve_ok = false; uint64_t last_iters = 0;
	PcgScalars *d_scal = nullptr, *h_scal = nullptr;
	double *partials = nullptr;     // [4][MAX_PARTIAL_BLOCKS]
	unsigned *ticket = nullptr;     // last-block counters
	uint32_t *ordinal = nullptr;    // exclusive scan of (cnt > 0) over local cells, ncl + 1 entries
	bool ordinal_valid = false;
	std::vector<MgLevel> mg;
	uint16_t *mg_mask = nullptr;   // level-0 coupling mask (mg.cu)
	std::vector<MgLevel> mg_agg;   // multi-GPU: agglomerated global coarse levels (experimental, mg.cu)
	int mg_agg_level = -1;         // distributed level they replace; -1 undecided, -2 none
	std::vector<int> mg_z0;        // global z of the first owned layer, per level (red-black parity)
	bool mg_valid = false;
	// one PCG iteration as a CUDA graph (pressure.cu)
	void *pcg_graph = nullptr;          // cudaGraphExec_t
	unsigned pcg_graph_launches = 0;    // kernels in it
	int pcg_graph_key = 0;              // configuration it was captured for

	// multi-GPU halos through peer memory (exchange.cu): one arena per rank, mapped by its z neighbours with CUDA IPC
	char *arena = nullptr;            // this rank's arena: header (flags, counters) + 2 x 2 receive slots
	char *arena_peer[2] = { nullptr, nullptr }; // [0] the upper neighbour's arena, [1] the lower one's (mapped)
	std::vector<char*> arena_all;     // every rank's arena as mapped here ([rank] = this rank's own)
	char **arena_all_d = nullptr;     // the same table in device memory
	size_t arena_slot = 0;            // bytes per receive slot
	unsigned long long halo_epoch = 0; // exchanges issued so far (identical on every rank)
	bool p2p = false;                 // arena mapped on both sides: halos bypass NCCL
	// multi-GPU particle exchange (exchange.cu)
	double *xsend[2] = { nullptr, nullptr }, *xrecv = nullptr; // [0] to / from the upper neighbour, [1] the lower one
	size_t xsend_cap[2] = { 0, 0 }, xrecv_cap = 0;             // in particles (15 doubles each)
	uint32_t *xcnt = nullptr; size_t xcnt_n = 0;               // per-block emigrant counts + their scans
	uint32_t *xcounts = nullptr, *h_xcounts = nullptr;         // the 4 message sizes (device / pinned host)

	// fluid sources (lfk_set_sources): entries = (cell, source) pairs sorted by cell, then by source order
	uint32_t src_entries = 0, src_count = 0; // entries in this rank's slab / sources
	bool src_active = false, src_coerce = false;
	uint32_t *src_cell = nullptr;    // [entries] local raw cell index
	uint32_t *src_gcell = nullptr;   // [entries] whole-grid raw cell index is src_gcell (RNG key) -- low 32 bits
	uint32_t *src_of = nullptr;      // [entries] source index
	uint32_t *src_need = nullptr;    // [entries + 1] particles to add per entry / their exclusive scan
	double *src_vel = nullptr;       // [sources][3]
	uint32_t *src_target = nullptr;  // [sources] target count (density cubed)
	uint16_t *src_map = nullptr;     // [ncl] 0, or 1 + index of the LAST coercing source that lists the cell
	uint64_t rng_seed = 0x5eed5eedull, rng_step = 0;

	// host transfers (transfer.cu)
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t xfer_ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // slot filled / slot free x 2, positions ready / copied
	void *pos_stage = nullptr; size_t pos_stage_bytes = 0; // device buffer of the asynchronous positions download
	bool pos_pending = false;

	// obstacle voxelisation (aux.cu): the voxel grid of the last lfk_voxelize_mesh
	uint8_t *vox = nullptr;
	long long vox_min[3] = { 0, 0, 0 };
	int vox_size[3] = { 0, 0, 0 };

	// scratch
	void *staging = nullptr; size_t staging_bytes = 0;
	uint32_t *scan_tmp = nullptr; size_t scan_tmp_n = 0;
	uint32_t *bigcells = nullptr; unsigned *bigcount = nullptr; unsigned bigcap = 0;
	double *d_reduce = nullptr;     // small device scratch for reductions (cfl etc.), 16 doubles
	// d_reduce[LFK_REDUCE_SPEED2]: max |v|^2 of the own particles as the last G2P left it; valid until anything else
	// writes particle velocities or changes the particle set (uploads, seeding, sources, checkpoint load)
	bool speed2_valid = false;
	double *h_reduce = nullptr;     // pinned

	// stats
	lfk_stats stats{};
	bool timing = false;
	int timer_depth = 0;
	cudaEvent_t ev[2] = { nullptr, nullptr };
};

#define LFK_MAX_PARTIAL_BLOCKS 16384

int lfk_fail(lfk_ctx *ctx, int code, const char *what, const char *file, int line);
// Small device -> pinned-host read-back issued by a KERNEL (the pinned buffer is device-accessible under unified
// addressing) instead of the device-to-host copy engine, whose queue is FIFO: behind an asynchronous positions download
// (transfer.cu) every cudaMemcpyAsync read-back of the step would wait for the whole 3 GB copy (measured: the step took
// 127 ms instead of 65).  Stream-ordered like the copy it replaces; the caller synchronises the stream as before.
int lfk_readback(lfk_ctx *c, void *pinned_host, const void *dev, size_t bytes);
const char *lfk_cuda_err_name(cudaError_t e);

#define LFK_CUDA(ctx, expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
	return lfk_fail((ctx), -(int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); } } while (0)
#define LFK_TRY(expr) do { int rc__ = (expr); if (rc__ != 0) { return rc__; } } while (0)
#define LFK_REQUIRE(ctx, cond, code, msg) do { if (!(cond)) { \
	return lfk_fail((ctx), (code), (msg), __FILE__, __LINE__); } } while (0)
// kernel launch + launch counter + launch-error check
#define LFK_LAUNCH(ctx, kernel, grid, block, smem, ...) do { \
	kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
	++(ctx)->stats.kernel_launches; \
	LFK_CUDA((ctx), cudaGetLastError()); } while (0)

static inline unsigned lfk_blocks(long long n, int block) {
	long long b = (n + block - 1) / block;
	return (unsigned)(b < 1 ? 1 : b);
}

struct PhaseTimer { // accumulates device time of a phase into stats.phase_ms when timing is enabled
	lfk_ctx *c; int phase; bool outer;
	PhaseTimer(lfk_ctx *ctx, int ph) : c(ctx), phase(ph), outer(false) { // nested timers: only the outermost counts
		if (c->timing && c->timer_depth++ == 0) {
			outer = true;
			cudaEventRecord(c->ev[0], c->stream);
		}
	}
	~PhaseTimer() {
		if (c->timing) { --c->timer_depth; }
		if (outer) {
			cudaEventRecord(c->ev[1], c->stream);
			cudaEventSynchronize(c->ev[1]);
			float ms = 0.f;
			cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
			c->stats.phase_ms[phase] += ms;
		}
	}
};

// ---- implemented in particles.cu ----
int lfkp_aos_to_soa(lfk_ctx *c, const void *d_aos, uint64_t n, uint64_t at = 0);   // particles [at, at + n) of the own view
int lfkp_soa_to_aos(lfk_ctx *c, void *d_aos, uint64_t n, uint64_t at = 0);
int lfkp_positions_to_aos(lfk_ctx *c, double *d_xyz, uint64_t n);
int lfkp_hash(lfk_ctx *c, bool lean);
int lfkp_materialise_vc(lfk_ctx *c);
int lfkp_permute_c(lfk_ctx *c);
int lfkp_advect(lfk_ctx *c, double dt);
int lfkp_collide(lfk_ctx *c);
int lfkp_advect_collide(lfk_ctx *c, double dt);       // fused, no old_position traffic
int lfkp_correct(lfk_ctx *c, double dt);
int lfkp_correct_collide(lfk_ctx *c, double dt);      // fused
int lfkp_g2p(lfk_ctx *c);
int lfkp_cfl(lfk_ctx *c, double *value);
int lfkp_seed_box(lfk_ctx *c, const double *start, const double *size, const double *vel, uint32_t dens,
	uint64_t seed, int append);
int lfkp_exclusive_scan_u32(lfk_ctx *c, const uint32_t *in, uint32_t *out, long long n, int from_flags);
int lfkp_coerce_sources(lfk_ctx *c);
int lfkp_update_sources(lfk_ctx *c, uint64_t *added);
int lfkp_reserve_particles(lfk_ctx *c, uint64_t n); // room for n entries behind `first` (may compact to first = 0)
static inline ParticleSoA lfk_own_view(const lfk_ctx *c) { // the own particles as arrays indexed from 0
	ParticleSoA v = c->P;
	for (int f = 0; f < PF_COUNT; ++f) { v.f[f] += c->first; }
	return v;
}

// ---- implemented in transfer.cu: chunked, double-buffered host transfers on a second stream; checkpoints ----
int lfk_reserve_staging(lfk_ctx *c, size_t bytes); // (lfk_api.cu)
int lfkt_upload_particles_pipelined(lfk_ctx *c, const void *aos152, uint64_t n);
int lfkt_download_particles_pipelined(lfk_ctx *c, void *aos152, uint64_t n);
int lfkt_destroy(lfk_ctx *c);

// ---- implemented in p2g.cu ----
int lfkg_p2g(lfk_ctx *c, double gravity_dt, bool add_gravity);
int lfkg_gravity(lfk_ctx *c, double dt);

// ---- implemented in pressure.cu ----
int lfks_build_system(lfk_ctx *c, double dt);
int lfks_solve(lfk_ctx *c, double dt, double *residual, uint64_t *iters, bool warm = false);
int lfks_apply_pressure(lfk_ctx *c, double dt);
int lfks_extrapolate(lfk_ctx *c);
int lfks_apply_a(lfk_ctx *c, double dt, const double *d_v_dense, double *d_out_dense);
int lfks_synthetic_projection(lfk_ctx *c, uint64_t seed);
int lfks_compact(lfk_ctx *c, const double *dense, double *d_out, const uint8_t *dense_u8, uint8_t *d_out_u8);
int lfks_expand(lfk_ctx *c, const double *d_compact, double *dense);
int lfks_ensure_ordinal(lfk_ctx *c);
int lfks_fluid_cells(lfk_ctx *c, uint64_t *d_out);
int lfks_export_flags(lfk_ctx *c, uint8_t *d_out_compact);

// ---- implemented in exchange.cu (multi-GPU; no-ops when nranks == 1) ----
int lfkx_init(lfk_ctx *c, const void *nccl_id128);
int lfkx_destroy(lfk_ctx *c);
int lfkx_check(lfk_ctx *c); // error if a peer-memory exchange timed out
int lfkx_halo_f64(lfk_ctx *c, double *field);          // fill both z ghost layers of a cell array from the neighbours
int lfkx_halo_f32(lfk_ctx *c, float *field, int nx, int ny, int nzl);
int lfkx_halo_u8(lfk_ctx *c, uint8_t *field);
int lfkx_layer_below(lfk_ctx *c, const double *field, double *dst); // dst <- the lower rank's layer z0 - 2 of field
// migrates particles that left the slab to the z neighbours and imports ghost copies of the neighbours' boundary
// layers; appends what arrives behind the own particles and returns the number of entries to sort
int lfkx_exchange_particles(lfk_ctx *c, uint64_t *n_in);
int lfkx_allreduce_sum(lfk_ctx *c, double *d_vals, int n);
int lfkx_allreduce_max(lfk_ctx *c, double *d_vals, int n);
int lfkx_allreduce_sum_f32(lfk_ctx *c, float *d_vals, int n);
// all-reduce of ONE PCG scalar (a field of *scal) over the ranks + its finaliser (FIN_*), in one kernel over peer memory
// when the arenas are mapped; returns 1 if it did, 0 if the caller has to use NCCL + k_finalize
int lfkx_allreduce_finalize(lfk_ctx *c, double *field, bool is_max, int which);

// device helpers ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// x / cell_size with the reference's rounding (a true IEEE division unless cell_size is a power of two, where the
// multiplication by the exact reciprocal gives the identical result at a fraction of the cost)
__device__ __forceinline__ double div_h(double x, const GridDesc &G) {
	return G.hpow2 ? x * G.inv_h : x / G.h;
}
__device__ __forceinline__ double dmax_std(double a, double b) { // std::max(a, b)
	return (a < b) ? b : a;
}
__device__ __forceinline__ double dclamp_std(double v, double lo, double hi) { // std::clamp
	return (v < lo) ? lo : (hi < v) ? hi : v;
}
// K1: cell coordinate of a position along one axis (reference src/simulation.cpp:251-261) -- must be bit-exact:
// IEEE subtraction, IEEE division, truncation toward zero, clamp to [0, n - 1]
__device__ __forceinline__ int cell_coord_clamped(double pos, double off, const GridDesc &G, int n) {
	double g = div_h(__dsub_rn(pos, off), G);
	g = dmax_std(g, 0.0);
	// static_cast<size_t>: truncation toward zero; cvt.rzi.u64.f64 saturates, and anything >= n clamps to n - 1
	unsigned long long v = (unsigned long long)g;
	return (int)(v < (unsigned long long)(n - 1) ? v : (unsigned long long)(n - 1));
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
	x ^= x >> 30;
	x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27;
	x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
// deterministic last-block reduction of per-block partials: every block stores its partial, the last block to
// arrive (ticket) reduces all of them in a fixed order.  Returns true in thread 0 of the last block.
__device__ __forceinline__ bool lfk_last_block(unsigned *ticket) {
	__shared__ bool is_last;
	__threadfence();
	if (threadIdx.x == 0) {
		unsigned t = atomicInc(ticket, gridDim.x - 1); // wraps back to 0 for the next use
		is_last = (t == gridDim.x - 1);
	}
	__syncthreads();
	return is_last;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v += __shfl_xor_sync(0xffffffffu, v, o);
	}
	return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	}
	return v;
}
// block-wide sum / max (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
	__shared__ double sm[32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	v = warp_sum(v);
	__syncthreads();
	if (lane == 0) { sm[w] = v; }
	__syncthreads();
	if (w == 0) {
		v = lane < (int)((blockDim.x + 31) >> 5) ? sm[lane] : 0.0;
		v = warp_sum(v);
	}
	return v;
}
__device__ __forceinline__ double block_max(double v) {
	__shared__ double sm[32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	v = warp_max(v);
	__syncthreads();
	if (lane == 0) { sm[w] = v; }
	__syncthreads();
	if (w == 0) {
		v = lane < (int)((blockDim.x + 31) >> 5) ? sm[lane] : -1.0e300;
		v = warp_max(v);
	}
	return v;
}

// ---- iteration over the owned cells without 64-bit div/mod: a warp takes one x-row at a time (lanes <-> 32
// consecutive cells: coalesced), rows are dealt round-robin to the warps of the grid.  f(x, y, lz, c) with c the
// local raw index.  Works for any 1-D block whose size is a multiple of 32 and any grid size.
template <typename F> __device__ __forceinline__ void for_own_cells(const GridDesc &G, F f) {
	const int wpb = (int)(blockDim.x >> 5), rows = G.ny * G.nzl;
	for (int row = (int)blockIdx.x * wpb + (int)(threadIdx.x >> 5); row < rows; row += (int)gridDim.x * wpb) {
		const int y = row % G.ny, lz = row / G.ny + 1;
		const long long base = (long long)G.nx * (y + (long long)G.ny * lz);
		for (int x = (int)(threadIdx.x & 31); x < G.nx; x += 32) {
			f(x, y, lz, base + x);
		}
	}
}
// Three-phase variant for the bandwidth-critical kernels: a warp takes a row and every lane handles up to U cells of
// it (x = xfirst + k * xstep).  `load` (nothing but loads into a plain struct R) runs for all U cells first, then
// `compute` (arithmetic only) and `store` -- so the loads of all U cells are in flight together instead of one
// dependent round trip per cell (these kernels are latency-bound otherwise: measured 15-25 % of L2 / DRAM
// throughput with one cell in flight per thread).
// colour < 0: all cells (xstep 32); colour 0 / 1: the cells with (x + y + z) & 1 == colour (xstep 64).
template <int U, typename R, typename FL, typename FC> __device__ __forceinline__ void rows_pipelined(int nx, int ny,
	int nzl, int zpar, int colour, FL load, FC compute_store) {
	const int wpb = (int)(blockDim.x >> 5), rows = ny * nzl, lane = (int)(threadIdx.x & 31);
	const int xstep = colour < 0 ? 32 : 64;
	for (int row = (int)blockIdx.x * wpb + (int)(threadIdx.x >> 5); row < rows; row += (int)gridDim.x * wpb) {
		const int y = row % ny, lz = row / ny + 1;
		const long long base = (long long)nx * (y + (long long)ny * lz);
		const int xfirst = colour < 0 ? lane : 2 * lane + ((y + (lz - 1 + zpar) + colour) & 1);
		for (int xw = 0; xw < nx; xw += U * xstep) { // warp-uniform trip count: every lane reaches the vote below
			const int xb = xw + xfirst;
			R raw[U];
			if (__all_sync(0xffffffffu, xb + (U - 1) * xstep < nx)) { // warp-uniform fast path: straight-line loads
#pragma unroll
				for (int k = 0; k < U; ++k) { raw[k] = load(xb + k * xstep, y, lz, base + xb + k * xstep); }
#pragma unroll
				for (int k = 0; k < U; ++k) { compute_store(xb + k * xstep, y, lz, base + xb + k * xstep, raw[k]); }
			} else {
#pragma unroll
				for (int k = 0; k < U; ++k) {
					const int x = xb + k * xstep;
					if (x < nx) { raw[k] = load(x, y, lz, base + x); }
				}
#pragma unroll
				for (int k = 0; k < U; ++k) {
					const int x = xb + k * xstep;
					if (x < nx) { compute_store(x, y, lz, base + x, raw[k]); }
				}
			}
		}
	}
}
// grid size for a row-per-warp kernel with `threads` threads per block
static inline unsigned lfk_row_blocks(const GridDesc &G, int threads, unsigned cap) {
	long long rows = (long long)G.ny * G.nzl, wpb = threads / 32;
	long long nb = (rows + wpb - 1) / wpb;
	if (nb < 1) { nb = 1; }
	return (unsigned)(nb > cap ? cap : nb);
}

// ---- pressure-system flag byte (one per cell): bits 0-2 non-solid neighbour count (the diagonal), bit 3 "is an
// unknown" (the cell holds particles, reference src/simulation.cpp:83-87), bit 4 type == fluid, bits 5-7
// type(+x / +y / +z neighbour) == fluid ---------------------------------------------------------------------------
#define FL_N(f) ((f) & 7u)
#define FL_L 8u
#define FL_SELF 16u
#define FL_XP 32u
#define FL_YP 64u
#define FL_ZP 128u

#define RED_BLOCKS 16384 // reduction kernels: one row per warp up to here, then rows are strided
#define RED_THREADS 256

// deterministic finish of a block-partial reduction (run by the last block): `op` 0 sum, 1 max
__device__ __forceinline__ double finish_partials(const double *partials, unsigned nblocks, int op) {
	double acc = op ? -1.0e300 : 0.0;
	for (unsigned k = threadIdx.x; k < nblocks; k += blockDim.x) {
		double t = partials[k];
		acc = op ? fmax(acc, t) : acc + t;
	}
	return op ? block_max(acc) : block_sum(acc);
}

// finalisers of the PCG scalars: run by the last block of the producing kernel on one GPU, or by k_finalize after
// the NCCL all-reduce of the local partial results on several GPUs
enum { FIN_BB = 0, FIN_ALPHA = 1, FIN_RESID = 2, FIN_BETA_FIRST = 3, FIN_BETA = 4 };
__device__ __forceinline__ void pcg_finalize(PcgScalars *scal, int which) {
	const double tolerance = scal->tolerance;
	switch (which) {
	case FIN_BB: // early-out of the reference (src/pressure_solver.cpp:29-35)
		scal->iters = 0;
		scal->resmax = 0.0;
		scal->done = scal->bb < 1e-6 ? 1 : 0;
		break;
	case FIN_ALPHA:
		scal->alpha = scal->sigma / scal->zs;
		break;
	case FIN_RESID: // :54-58 (two-sided: max |r|, which implies the reference's one-sided max r < tolerance)
		scal->iters += 1;
		if (scal->resmax < tolerance) { scal->done = 1; }
		break;
	case FIN_BETA_FIRST: // :38-42
		scal->sigma = scal->sigma_new;
		scal->beta = 0.0;
		break;
	default: // :62-68
		scal->beta = scal->sigma_new / scal->sigma;
		scal->sigma = scal->sigma_new;
		break;
	}
}
#endif
