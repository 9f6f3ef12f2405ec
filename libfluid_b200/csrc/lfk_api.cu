// The C ABI of include/lfk.h: context lifetime, host<->device state transfer, stage entry points and the fused
// time step.  Everything here is host code; the kernels live in particles.cu / p2g.cu / pressure.cu / mg.cu.
#include "lfk_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

static thread_local std::string g_create_error;

int lfk_fail(lfk_ctx *ctx, int code, const char *what, const char *file, int line) {
	char buf[512];
	const char *base = strrchr(file, '/');
	snprintf(buf, sizeof(buf), "lfk error %d: %s (%s:%d)", code, what, base ? base + 1 : file, line);
	if (ctx) {
		ctx->err = buf;
	} else {
		g_create_error = buf;
	}
	return code;
}

__global__ void k_readback(unsigned *__restrict__ dst, const unsigned *__restrict__ src, unsigned nwords) {
	for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) { dst[i] = src[i]; }
	__threadfence_system();
}
int lfk_readback(lfk_ctx *c, void *pinned_host, const void *dev, size_t bytes) {
	LFK_REQUIRE(c, bytes % 4 == 0 && bytes <= 4096, LFK_E_INVALID, "lfk_readback: small word-sized blocks only");
	LFK_LAUNCH(c, k_readback, 1, 64, 0, (unsigned*)pinned_host, (const unsigned*)dev, (unsigned)(bytes / 4));
	return 0;
}

extern "C" int lfk_abi_version(void) {
	return LFK_ABI_VERSION;
}

extern "C" const char *lfk_last_error(const lfk_ctx *ctx) {
	return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

// ---- allocation helpers -------------------------------------------------------------------------------------
template <typename T> static int dev_alloc(lfk_ctx *c, T **p, size_t n) {
	*p = nullptr;
	LFK_CUDA(c, cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)));
	return 0;
}
template <typename T> static void dev_free(T *&p) {
	if (p) {
		cudaFree(p);
		p = nullptr;
	}
}

int lfkp_reserve_particles(lfk_ctx *c, uint64_t n) {
	if (c->first + n <= c->cap) { return 0; }
	if (c->np > 0) { LFK_TRY(lfkp_materialise_vc(c)); } // the permutation buffer is reallocated below
	uint64_t ncap = std::max<uint64_t>(n, c->cap + c->cap / 4);
	// multi-GPU: every sort appends the neighbours' boundary layers (ghost copies) and immigrants behind the own
	// particles; room for two layers per side at 8 particles per cell up front avoids a reallocation in the first step
	if (c->nranks > 1) { ncap += (uint64_t)c->g.sxy * 32; }
	ncap = (ncap + 1023) / 1024 * 1024;
	ParticleSoA nP{}, nA{};
	uint32_t *nkey = nullptr, *nkalt = nullptr, *nslot = nullptr, *nperm = nullptr;
	for (int f = 0; f < PF_COUNT; ++f) {
		LFK_TRY(dev_alloc(c, &nP.f[f], ncap));
		LFK_TRY(dev_alloc(c, &nA.f[f], ncap));
	}
	LFK_TRY(dev_alloc(c, &nkey, ncap));
	LFK_TRY(dev_alloc(c, &nkalt, ncap));
	LFK_TRY(dev_alloc(c, &nslot, ncap));
	LFK_TRY(dev_alloc(c, &nperm, ncap));
	if (c->np > 0) { // keep the own particles, now from index 0 (the cell table, if any, becomes invalid)
		int nf = c->old_valid ? PF_COUNT : 15;
		for (int f = 0; f < nf; ++f) {
			LFK_CUDA(c, cudaMemcpyAsync(nP.f[f], c->P.f[f] + c->first, c->np * sizeof(double), cudaMemcpyDeviceToDevice,
				c->stream));
		}
		LFK_CUDA(c, cudaMemcpyAsync(nkey, c->key + c->first, c->np * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
		LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	if (c->first != 0) { c->table_valid = false; }
	c->first = 0;
	c->ntot = c->np;
	for (int f = 0; f < PF_COUNT; ++f) {
		dev_free(c->P.f[f]);
		dev_free(c->Palt.f[f]);
	}
	dev_free(c->key);
	dev_free(c->key_alt);
	dev_free(c->slot);
	dev_free(c->perm);
	c->P = nP;
	c->Palt = nA;
	c->key = nkey;
	c->key_alt = nkalt;
	c->slot = nslot;
	c->perm = nperm;
	c->cap = ncap;
	return 0;
}

int lfk_reserve_staging(lfk_ctx *c, size_t bytes) {
	if (bytes <= c->staging_bytes) { return 0; }
	if (c->staging) {
		cudaFree(c->staging);
		c->staging = nullptr;
		c->staging_bytes = 0;
	}
	LFK_CUDA(c, cudaMalloc(&c->staging, bytes));
	c->staging_bytes = bytes;
	return 0;
}

// cell-centre coordinates exactly as the reference builds them: start at offset + h/2 and ADD h per cell
// (src/simulation.cpp:294-300, 347-353) -- not (i + 0.5) * h, which rounds differently for h != 1.
static int upload_centres(lfk_ctx *c) {
	const GridDesc &G = c->g;
	const int size[3] = { G.nx, G.ny, G.nz };
	for (int d = 0; d < 3; ++d) {
		std::vector<double> t((size_t)size[d]);
		double pos = G.off[d] + 0.5 * G.h;
		for (int i = 0; i < size[d]; ++i, pos += G.h) {
			t[(size_t)i] = pos;
		}
		LFK_CUDA(c, cudaMemcpyAsync(c->ctr[d], t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	return 0;
}

__global__ void k_init_types(GridDesc G, uint8_t *typ) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= G.ncl) { return; }
	int lz = (int)(i / G.sxy);
	int z = lz - 1 + G.z0;
	// ghost layers outside the domain are "solid"; everything else starts as air (mac_grid.h:26)
	typ[i] = (z < 0 || z >= G.nz) ? LFK_CELL_SOLID : LFK_CELL_AIR;
}

static void default_params(lfk_params *p) {
	memset(p, 0, sizeof(*p));
	p->cell_size = NAN;
	p->density = 1.0;
	p->boundary_skin_width = 0.1;
	p->correction_stiffness = 5.0;
	p->blending_factor = 1.0;
	p->cfl_number = 3.0;
	p->tolerance = 1e-6;
	p->method = LFK_METHOD_APIC;
	p->extrapolation_iterations = 1;
	p->max_iterations = 200;
	p->preconditioner = LFK_PRECOND_MULTIGRID;
}

static int free_all(lfk_ctx *c) {
	for (int f = 0; f < PF_COUNT; ++f) {
		dev_free(c->P.f[f]);
		dev_free(c->Palt.f[f]);
	}
	dev_free(c->key);
	dev_free(c->key_alt);
	dev_free(c->slot);
	dev_free(c->perm);
	for (int d = 0; d < 3; ++d) {
		dev_free(c->vel[d]);
		dev_free(c->vel_old[d]);
		dev_free(c->ctr[d]);
	}
	dev_free(c->typ);
	dev_free(c->cnt);
	dev_free(c->begin);
	dev_free(c->valid[0]);
	dev_free(c->valid[1]);
	dev_free(c->wlow[0]);
	dev_free(c->wlow[1]);
	dev_free(c->flags);
	dev_free(c->b);
	dev_free(c->p);
	dev_free(c->r);
	dev_free(c->z);
	dev_free(c->s);
	dev_free(c->d_scal);
	dev_free(c->partials);
	dev_free(c->ticket);
	dev_free(c->ordinal);
	dev_free(c->scan_tmp);
	dev_free(c->bigcells);
	dev_free(c->bigcount);
	dev_free(c->d_reduce);
	dev_free(c->xsend[0]);
	dev_free(c->xsend[1]);
	dev_free(c->xrecv);
	dev_free(c->xcnt);
	dev_free(c->src_cell);
	dev_free(c->src_gcell);
	dev_free(c->src_of);
	dev_free(c->src_need);
	dev_free(c->src_vel);
	dev_free(c->src_target);
	dev_free(c->src_map);
	dev_free(c->vox);
	dev_free(c->xcounts);
	if (c->h_xcounts) { cudaFreeHost(c->h_xcounts); c->h_xcounts = nullptr; }
	if (c->staging) { cudaFree(c->staging); c->staging = nullptr; }
	if (c->h_scal) { cudaFreeHost(c->h_scal); c->h_scal = nullptr; }
	if (c->h_reduce) { cudaFreeHost(c->h_reduce); c->h_reduce = nullptr; }
	return 0;
}

int lfkm_free(lfk_ctx *c); // mg.cu
int lfks_free_graph(lfk_ctx *c); // pressure.cu

extern "C" int lfk_set_tuning(lfk_ctx *c, const char *key, int value);

extern "C" int lfk_create(lfk_ctx **out, uint64_t nx, uint64_t ny, uint64_t nz, int device, void *stream,
	int nranks, int rank, const void *nccl_id128) {
	if (!out) { return lfk_fail(nullptr, LFK_E_INVALID, "out is NULL", __FILE__, __LINE__); }
	*out = nullptr;
	if (nx < 1 || ny < 1 || nz < 1 || nx > 4096 || ny > 4096 || nz > 4096) {
		return lfk_fail(nullptr, LFK_E_INVALID, "grid size out of range", __FILE__, __LINE__);
	}
	// a particle moves at most cfl_number (3) cells per step: with slabs >= 4 cells thick an immigrant cannot reach
	// the far boundary layer of its new slab within the step it arrives in (one exchange per step suffices)
	if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && nz < 4 * (uint64_t)nranks)) {
		return lfk_fail(nullptr, LFK_E_INVALID, "bad rank layout (slabs must be >= 4 cells thick)", __FILE__, __LINE__);
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1) { // NO CPU fallback
		return lfk_fail(nullptr, LFK_E_NO_DEVICE, "no CUDA device available; lfk has no CPU fallback", __FILE__, __LINE__);
	}
	if (device < 0 || device >= ndev) {
		return lfk_fail(nullptr, LFK_E_INVALID, "device ordinal out of range", __FILE__, __LINE__);
	}
	lfk_ctx *c = new (std::nothrow) lfk_ctx();
	if (!c) { return lfk_fail(nullptr, LFK_E_INVALID, "out of host memory", __FILE__, __LINE__); }
	c->device = device;
	c->nranks = nranks;
	c->rank = rank;
#define CREATE_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
	int rc__ = lfk_fail(nullptr, -(int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
	free_all(c); delete c; return rc__; } } while (0)
#define CREATE_TRY(expr) do { int rc__ = (expr); if (rc__ != 0) { g_create_error = c->err; free_all(c); delete c; return rc__; } } while (0)
	CREATE_CUDA(cudaSetDevice(device));
	if (stream) {
		c->stream = (cudaStream_t)stream;
	} else {
		CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		c->own_stream = true;
	}
	default_params(&c->prm);
	GridDesc &G = c->g;
	G.nx = (int)nx;
	G.ny = (int)ny;
	G.nz = (int)nz;
	// contiguous z-slabs, remainder spread over the first ranks
	int base = (int)(nz / nranks), rem = (int)(nz % nranks);
	G.z0 = rank * base + std::min(rank, rem);
	G.nzl = base + (rank < rem ? 1 : 0);
	G.nlz = G.nzl + 2;
	G.sxy = (long long)nx * ny;
	G.ncl = G.sxy * G.nlz;
	G.nown = G.sxy * G.nzl;
	G.h = NAN;
	G.inv_h = NAN;
	G.hpow2 = 0;
	G.off[0] = G.off[1] = G.off[2] = 0.0;
	size_t ncl = (size_t)G.ncl;
	for (int d = 0; d < 3; ++d) {
		CREATE_TRY(dev_alloc(c, &c->vel[d], ncl));
		CREATE_CUDA(cudaMemsetAsync(c->vel[d], 0, ncl * sizeof(double), c->stream));
	}
	CREATE_TRY(dev_alloc(c, &c->ctr[0], (size_t)nx));
	CREATE_TRY(dev_alloc(c, &c->ctr[1], (size_t)ny));
	CREATE_TRY(dev_alloc(c, &c->ctr[2], (size_t)nz));
	CREATE_TRY(dev_alloc(c, &c->typ, ncl));
	CREATE_TRY(dev_alloc(c, &c->cnt, ncl + 1));   // + the graveyard bin of the cell sort
	CREATE_TRY(dev_alloc(c, &c->begin, ncl + 2));
	CREATE_TRY(dev_alloc(c, &c->valid[0], ncl));
	CREATE_TRY(dev_alloc(c, &c->valid[1], ncl));
	if (nranks > 1) {
		for (int k = 0; k < 2; ++k) {
			CREATE_TRY(dev_alloc(c, &c->wlow[k], (size_t)G.sxy));
			CREATE_CUDA(cudaMemsetAsync(c->wlow[k], 0, (size_t)G.sxy * sizeof(double), c->stream));
		}
	}
	CREATE_TRY(dev_alloc(c, &c->flags, ncl));
	CREATE_TRY(dev_alloc(c, &c->b, ncl));
	CREATE_TRY(dev_alloc(c, &c->p, ncl));
	CREATE_TRY(dev_alloc(c, &c->r, ncl));
	CREATE_TRY(dev_alloc(c, &c->z, ncl));
	CREATE_TRY(dev_alloc(c, &c->s, ncl));
	CREATE_TRY(dev_alloc(c, &c->ordinal, ncl + 1));
	CREATE_TRY(dev_alloc(c, &c->d_scal, 1));
	CREATE_TRY(dev_alloc(c, &c->partials, (size_t)LFK_MAX_PARTIAL_BLOCKS));
	CREATE_TRY(dev_alloc(c, &c->ticket, 4));
	c->bigcap = (unsigned)std::min<size_t>(ncl, (size_t)1 << 22);
	CREATE_TRY(dev_alloc(c, &c->bigcells, (size_t)c->bigcap));
	CREATE_TRY(dev_alloc(c, &c->bigcount, 1));
	CREATE_TRY(dev_alloc(c, &c->d_reduce, 16));
	CREATE_CUDA(cudaMallocHost((void**)&c->h_scal, sizeof(PcgScalars)));
	CREATE_CUDA(cudaMallocHost((void**)&c->h_reduce, 16 * sizeof(double)));
	CREATE_TRY(dev_alloc(c, &c->xcounts, 8));
	CREATE_CUDA(cudaMallocHost((void**)&c->h_xcounts, 8 * sizeof(uint32_t)));
	CREATE_CUDA(cudaMemsetAsync(c->cnt, 0, (ncl + 1) * sizeof(uint32_t), c->stream));
	CREATE_CUDA(cudaMemsetAsync(c->begin, 0, (ncl + 2) * sizeof(uint32_t), c->stream));
	CREATE_CUDA(cudaMemsetAsync(c->flags, 0, ncl, c->stream));
	CREATE_CUDA(cudaMemsetAsync(c->ticket, 0, 4 * sizeof(unsigned), c->stream));
	CREATE_CUDA(cudaMemsetAsync(c->d_scal, 0, sizeof(PcgScalars), c->stream));
	double *zero_these[] = { c->b, c->p, c->r, c->z, c->s };
	for (double *ptr : zero_these) {
		CREATE_CUDA(cudaMemsetAsync(ptr, 0, ncl * sizeof(double), c->stream));
	}
	k_init_types<<<lfk_blocks(G.ncl, 256), 256, 0, c->stream>>>(G, c->typ);
	CREATE_CUDA(cudaGetLastError());
	CREATE_CUDA(cudaEventCreate(&c->ev[0]));
	CREATE_CUDA(cudaEventCreate(&c->ev[1]));
	CREATE_TRY(lfkx_init(c, nccl_id128));
	CREATE_CUDA(cudaStreamSynchronize(c->stream));
#undef CREATE_CUDA
#undef CREATE_TRY
	// LFK_TUNE="key=value,key=value": lfk_set_tuning for every context of the process (A/B runs of unmodified callers)
	if (const char *env = getenv("LFK_TUNE")) {
		std::string spec(env);
		size_t pos = 0;
		while (pos < spec.size()) {
			size_t end = spec.find(',', pos);
			if (end == std::string::npos) { end = spec.size(); }
			const std::string item = spec.substr(pos, end - pos);
			const size_t eq = item.find('=');
			if (eq != std::string::npos) { lfk_set_tuning(c, item.substr(0, eq).c_str(), atoi(item.c_str() + eq + 1)); }
			pos = end + 1;
		}
	}
	*out = c;
	return 0;
}

extern "C" int lfk_destroy(lfk_ctx *c) {
	if (!c) { return 0; }
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	lfks_free_graph(c);
	lfkt_destroy(c);
	lfkx_destroy(c);
	lfkm_free(c);
	free_all(c);
	if (c->ev[0]) { cudaEventDestroy(c->ev[0]); }
	if (c->ev[1]) { cudaEventDestroy(c->ev[1]); }
	if (c->own_stream) { cudaStreamDestroy(c->stream); }
	delete c;
	return 0;
}

extern "C" int lfk_set_params(lfk_ctx *c, const lfk_params *p) {
	if (!c || !p) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, p->cell_size > 0.0 && std::isfinite(p->cell_size), LFK_E_INVALID, "cell_size must be set (> 0)");
	LFK_REQUIRE(c, p->method >= 0 && p->method <= 2, LFK_E_INVALID, "unknown simulation method");
	LFK_REQUIRE(c, p->preconditioner >= 0 && p->preconditioner <= 1, LFK_E_INVALID, "unknown preconditioner");
	LFK_REQUIRE(c, p->density > 0.0, LFK_E_INVALID, "density must be positive");
	LFK_REQUIRE(c, p->max_iterations >= 0 && p->extrapolation_iterations >= 0, LFK_E_INVALID, "negative count");
	bool geom = c->g.h != p->cell_size || c->g.off[0] != p->grid_offset[0] || c->g.off[1] != p->grid_offset[1] ||
		c->g.off[2] != p->grid_offset[2] || std::isnan(c->g.h);
	c->prm = *p;
	c->g.h = p->cell_size;
	c->g.inv_h = 1.0 / p->cell_size;
	{
		int ex = 0;
		c->g.hpow2 = std::frexp(p->cell_size, &ex) == 0.5 ? 1 : 0;
	}
	for (int d = 0; d < 3; ++d) {
		c->g.off[d] = p->grid_offset[d];
	}
	if (p->method == LFK_METHOD_FLIP && !c->vel_old[0]) {
		for (int d = 0; d < 3; ++d) {
			LFK_TRY(dev_alloc(c, &c->vel_old[d], (size_t)c->g.ncl));
			LFK_CUDA(c, cudaMemsetAsync(c->vel_old[d], 0, (size_t)c->g.ncl * sizeof(double), c->stream));
		}
	}
	if (geom) {
		LFK_TRY(upload_centres(c));
		c->table_valid = false;
	}
	c->system_valid = false;
	return 0;
}

extern "C" int lfk_get_params(const lfk_ctx *c, lfk_params *p) {
	if (!c || !p) { return LFK_E_INVALID; }
	*p = c->prm;
	return 0;
}

extern "C" int lfk_sync(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_slab(const lfk_ctx *c, uint64_t *z_begin, uint64_t *z_end) {
	if (!c) { return LFK_E_INVALID; }
	if (z_begin) { *z_begin = (uint64_t)c->g.z0; }
	if (z_end) { *z_end = (uint64_t)(c->g.z0 + c->g.nzl); }
	return 0;
}

#define NEED_PARAMS(c) LFK_REQUIRE(c, std::isfinite((c)->g.h), LFK_E_STATE, "lfk_set_params has not been called")

// ---- particles ----------------------------------------------------------------------------------------------
extern "C" int lfk_upload_particles(lfk_ctx *c, const void *aos152, uint64_t n) {
	if (!c || (!aos152 && n)) { return LFK_E_INVALID; }
	PhaseTimer T(c, LFK_PHASE_TRANSFER);
	LFK_CUDA(c, cudaSetDevice(c->device));
	c->np = 0;
	c->first = 0;
	c->ntot = 0;
	c->v_deferred = false;
	c->c_deferred = false;
	c->speed2_valid = false;
	LFK_TRY(lfkp_reserve_particles(c, n));
	if (n > 0) {
		c->np = n;
		LFK_TRY(lfkt_upload_particles_pipelined(c, aos152, n));
	}
	c->np = n;
	c->ntot = n;
	c->old_valid = true;
	c->table_valid = false;
	c->keys_valid = true;
	c->v_deferred = false;
	c->c_deferred = false;
	return 0;
}

extern "C" int lfk_num_particles(lfk_ctx *c, uint64_t *n) {
	if (!c || !n) { return LFK_E_INVALID; }
	*n = c->np;
	return 0;
}

extern "C" int lfk_download_particles(lfk_ctx *c, void *aos152, uint64_t capacity, uint64_t *n) {
	if (!c) { return LFK_E_INVALID; }
	PhaseTimer T(c, LFK_PHASE_TRANSFER);
	if (n) { *n = c->np; }
	LFK_REQUIRE(c, capacity >= c->np, LFK_E_CAPACITY, "particle buffer too small");
	if (c->np == 0) { return 0; }
	LFK_REQUIRE(c, aos152 != nullptr, LFK_E_INVALID, "NULL particle buffer");
	LFK_TRY(lfkp_materialise_vc(c));
	return lfkt_download_particles_pipelined(c, aos152, c->np);
}

extern "C" int lfk_download_positions(lfk_ctx *c, double *xyz, uint64_t capacity, uint64_t *n) {
	if (!c) { return LFK_E_INVALID; }
	PhaseTimer T(c, LFK_PHASE_TRANSFER);
	if (n) { *n = c->np; }
	LFK_REQUIRE(c, capacity >= c->np, LFK_E_CAPACITY, "position buffer too small");
	if (c->np == 0) { return 0; }
	LFK_REQUIRE(c, xyz != nullptr, LFK_E_INVALID, "NULL position buffer");
	LFK_TRY(lfk_reserve_staging(c, (size_t)c->np * 24));
	LFK_TRY(lfkp_positions_to_aos(c, (double*)c->staging, c->np));
	LFK_CUDA(c, cudaMemcpyAsync(xyz, c->staging, (size_t)c->np * 24, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

// ---- cells (32-byte AoS: 3 doubles + type byte + 7 padding bytes) ---------------------------------------------
__global__ void k_cells_from_aos(GridDesc G, const unsigned long long *__restrict__ aos, double *__restrict__ u,
	double *__restrict__ v, double *__restrict__ w, uint8_t *__restrict__ typ) {
	// aos points at the first cell of layer max(z0 - 1, 0); fills owned layers and in-domain ghost layers
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= G.ncl) { return; }
	int lz = (int)(i / G.sxy);
	int z = lz - 1 + G.z0;
	if (z < 0 || z >= G.nz) { return; }
	int zfirst = G.z0 - 1 < 0 ? 0 : G.z0 - 1;
	long long src = (i - (long long)lz * G.sxy) + (long long)(z - zfirst) * G.sxy;
	const unsigned long long *rec = aos + 4 * src;
	u[i] = __longlong_as_double((long long)rec[0]);
	v[i] = __longlong_as_double((long long)rec[1]);
	w[i] = __longlong_as_double((long long)rec[2]);
	if (typ) { typ[i] = (uint8_t)(rec[3] & 0xffull); }
}
__global__ void k_cells_to_aos(GridDesc G, unsigned long long *__restrict__ aos, const double *__restrict__ u,
	const double *__restrict__ v, const double *__restrict__ w, const uint8_t *__restrict__ typ) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x; // owned cells only
	if (own >= G.nown) { return; }
	long long i = own + G.sxy;
	unsigned long long *rec = aos + 4 * own;
	rec[0] = (unsigned long long)__double_as_longlong(u[i]);
	rec[1] = (unsigned long long)__double_as_longlong(v[i]);
	rec[2] = (unsigned long long)__double_as_longlong(w[i]);
	rec[3] = (unsigned long long)typ[i];
}

// whole_grid: aos32 is the whole nx*ny*nz array; otherwise it starts at layer max(z0 - 1, 0) (slab + in-domain ghosts)
static int upload_cells_impl(lfk_ctx *c, const void *aos32, double **vel, uint8_t *typ, bool whole_grid = true) {
	const GridDesc &G = c->g;
	PhaseTimer T(c, LFK_PHASE_TRANSFER);
	int zfirst = std::max(G.z0 - 1, 0), zlast = std::min(G.z0 + G.nzl + 1, G.nz);
	size_t ncopy = (size_t)(zlast - zfirst) * (size_t)G.sxy;
	LFK_TRY(lfk_reserve_staging(c, ncopy * 32));
	LFK_CUDA(c, cudaMemcpyAsync(c->staging, (const char*)aos32 + (whole_grid ? (size_t)zfirst * G.sxy * 32 : 0), ncopy * 32,
		cudaMemcpyHostToDevice, c->stream));
	LFK_LAUNCH(c, k_cells_from_aos, lfk_blocks(G.ncl, 256), 256, 0, G, (const unsigned long long*)c->staging,
		vel[0], vel[1], vel[2], typ);
	return 0;
}
static int download_cells_impl(lfk_ctx *c, void *aos32, double **vel, bool whole_grid = true) {
	const GridDesc &G = c->g;
	PhaseTimer T(c, LFK_PHASE_TRANSFER);
	size_t nown = (size_t)G.nown;
	LFK_TRY(lfk_reserve_staging(c, nown * 32));
	LFK_LAUNCH(c, k_cells_to_aos, lfk_blocks(G.nown, 256), 256, 0, G, (unsigned long long*)c->staging, vel[0],
		vel[1], vel[2], c->typ);
	LFK_CUDA(c, cudaMemcpyAsync((char*)aos32 + (whole_grid ? (size_t)G.z0 * G.sxy * 32 : 0), c->staging, nown * 32,
		cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_upload_cells(lfk_ctx *c, const void *aos32) {
	if (!c || !aos32) { return LFK_E_INVALID; }
	LFK_TRY(upload_cells_impl(c, aos32, c->vel, c->typ));
	c->system_valid = false;
	c->pressure_valid = false;
	return 0;
}
extern "C" int lfk_download_cells(lfk_ctx *c, void *aos32) {
	if (!c || !aos32) { return LFK_E_INVALID; }
	return download_cells_impl(c, aos32, c->vel);
}
extern "C" int lfk_upload_cells_slab(lfk_ctx *c, const void *aos32_slab) {
	if (!c || !aos32_slab) { return LFK_E_INVALID; }
	LFK_TRY(upload_cells_impl(c, aos32_slab, c->vel, c->typ, false));
	c->system_valid = false;
	c->pressure_valid = false;
	return 0;
}
extern "C" int lfk_download_cells_slab(lfk_ctx *c, void *aos32_own) {
	if (!c || !aos32_own) { return LFK_E_INVALID; }
	return download_cells_impl(c, aos32_own, c->vel, false);
}
extern "C" int lfk_upload_old_cells(lfk_ctx *c, const void *aos32) {
	if (!c || !aos32) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->vel_old[0] != nullptr, LFK_E_STATE, "old grid exists only for LFK_METHOD_FLIP");
	return upload_cells_impl(c, aos32, c->vel_old, nullptr);
}
extern "C" int lfk_download_old_cells(lfk_ctx *c, void *aos32) {
	if (!c || !aos32) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->vel_old[0] != nullptr, LFK_E_STATE, "old grid exists only for LFK_METHOD_FLIP");
	return download_cells_impl(c, aos32, c->vel_old);
}

__global__ void k_table_to_u64(GridDesc G, const uint32_t *__restrict__ begin, unsigned long long *__restrict__ out_b,
	unsigned long long *__restrict__ out_c, uint32_t first) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long i = own + G.sxy;
	uint32_t b = begin[i], n = begin[i + 1] - b;
	out_c[own] = n;
	out_b[own] = n ? b - first : 0; // the reference leaves begin = 0 in empty cells (reset_space_hash)
}

extern "C" int lfk_download_table(lfk_ctx *c, uint64_t *begin, uint64_t *count) {
	if (!c || !begin || !count) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->table_valid, LFK_E_STATE, "no valid cell table (call lfk_hash)");
	const GridDesc &G = c->g;
	size_t nown = (size_t)G.nown;
	LFK_TRY(lfk_reserve_staging(c, nown * 16));
	unsigned long long *db = (unsigned long long*)c->staging, *dc = db + nown;
	LFK_LAUNCH(c, k_table_to_u64, lfk_blocks(G.nown, 256), 256, 0, G, c->begin, db, dc, (uint32_t)c->first);
	size_t off = (size_t)G.z0 * G.sxy;
	LFK_CUDA(c, cudaMemcpyAsync(begin + off, db, nown * 8, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaMemcpyAsync(count + off, dc, nown * 8, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

static int fetch_num_fluid(lfk_ctx *c, uint64_t *nf) {
	LFK_TRY(lfks_ensure_ordinal(c));
	uint32_t v = 0;
	LFK_CUDA(c, cudaMemcpyAsync(&v, c->ordinal + c->g.ncl, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	*nf = v;
	c->stats.num_fluid_cells = v;
	return 0;
}

extern "C" int lfk_num_fluid_cells(lfk_ctx *c, uint64_t *nf) {
	if (!c || !nf) { return LFK_E_INVALID; }
	return fetch_num_fluid(c, nf);
}

extern "C" int lfk_download_fluid_cells(lfk_ctx *c, uint64_t *raw, uint64_t capacity) {
	if (!c) { return LFK_E_INVALID; }
	uint64_t nf = 0;
	LFK_TRY(fetch_num_fluid(c, &nf));
	LFK_REQUIRE(c, capacity >= nf, LFK_E_CAPACITY, "fluid-cell buffer too small");
	if (nf == 0) { return 0; } // an empty list may come with a NULL buffer (std::vector::data())
	LFK_REQUIRE(c, raw != nullptr, LFK_E_INVALID, "NULL fluid-cell buffer");
	LFK_TRY(lfk_reserve_staging(c, (size_t)nf * 8));
	LFK_TRY(lfks_fluid_cells(c, (uint64_t*)c->staging));
	LFK_CUDA(c, cudaMemcpyAsync(raw, c->staging, (size_t)nf * 8, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

// ---- stages -------------------------------------------------------------------------------------------------
extern "C" int lfk_hash(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_hash(c, false);
}
extern "C" int lfk_advect(lfk_ctx *c, double dt) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_advect(c, dt);
}
extern "C" int lfk_collide(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_collide(c);
}
extern "C" int lfk_p2g(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkg_p2g(c, 0.0, false);
}
extern "C" int lfk_gravity(lfk_ctx *c, double dt) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkg_gravity(c, dt);
}
extern "C" int lfk_pressure_solve(lfk_ctx *c, double dt, double *residual, uint64_t *iterations) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfks_solve(c, dt, residual, iterations);
}

extern "C" int lfk_download_rhs(lfk_ctx *c, double dt, double *b, uint8_t *flags, uint64_t capacity) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	if (!c->system_valid || c->system_dt != dt) {
		LFK_TRY(lfks_build_system(c, dt));
	}
	uint64_t nf = 0;
	LFK_TRY(fetch_num_fluid(c, &nf));
	LFK_REQUIRE(c, capacity >= nf, LFK_E_CAPACITY, "rhs buffer too small");
	if (nf == 0) { return 0; }
	LFK_TRY(lfk_reserve_staging(c, (size_t)nf * 9));
	double *db = (double*)c->staging;
	uint8_t *df = (uint8_t*)(db + nf);
	LFK_TRY(lfks_compact(c, c->b, db, c->flags, df));
	if (b) { LFK_CUDA(c, cudaMemcpyAsync(b, db, (size_t)nf * 8, cudaMemcpyDeviceToHost, c->stream)); }
	if (flags) { LFK_CUDA(c, cudaMemcpyAsync(flags, df, (size_t)nf, cudaMemcpyDeviceToHost, c->stream)); }
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_download_pressure(lfk_ctx *c, double *p, uint64_t capacity) {
	if (!c) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->pressure_valid, LFK_E_STATE, "no pressure available (call lfk_pressure_solve)");
	uint64_t nf = 0;
	LFK_TRY(fetch_num_fluid(c, &nf));
	LFK_REQUIRE(c, capacity >= nf, LFK_E_CAPACITY, "pressure buffer too small");
	if (nf == 0) { return 0; }
	LFK_REQUIRE(c, p != nullptr, LFK_E_INVALID, "NULL pressure buffer");
	LFK_TRY(lfk_reserve_staging(c, (size_t)nf * 8));
	LFK_TRY(lfks_compact(c, c->p, (double*)c->staging, nullptr, nullptr));
	LFK_CUDA(c, cudaMemcpyAsync(p, c->staging, (size_t)nf * 8, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_upload_pressure(lfk_ctx *c, const double *p, uint64_t n) {
	if (!c || (!p && n)) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->system_valid, LFK_E_STATE, "build the system first (lfk_download_rhs / lfk_pressure_solve)");
	uint64_t nf = 0;
	LFK_TRY(fetch_num_fluid(c, &nf));
	LFK_REQUIRE(c, n == nf, LFK_E_INVALID, "pressure vector length != number of fluid cells");
	LFK_TRY(lfk_reserve_staging(c, (size_t)(nf ? nf : 1) * 8));
	if (nf) { LFK_CUDA(c, cudaMemcpyAsync(c->staging, p, (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream)); }
	LFK_TRY(lfks_expand(c, (const double*)c->staging, c->p));
	c->pressure_valid = true;
	return 0;
}

extern "C" int lfk_apply_a(lfk_ctx *c, double dt, const double *v, double *out, uint64_t n) {
	if (!c || !v || !out) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	if (!c->system_valid || c->system_dt != dt) {
		LFK_TRY(lfks_build_system(c, dt));
	}
	uint64_t nf = 0;
	LFK_TRY(fetch_num_fluid(c, &nf));
	LFK_REQUIRE(c, n == nf, LFK_E_INVALID, "vector length != number of fluid cells");
	if (nf == 0) { return 0; }
	LFK_TRY(lfk_reserve_staging(c, (size_t)nf * 8));
	LFK_CUDA(c, cudaMemcpyAsync(c->staging, v, (size_t)nf * 8, cudaMemcpyHostToDevice, c->stream));
	LFK_TRY(lfks_expand(c, (const double*)c->staging, c->s));
	if (c->nranks > 1) { LFK_TRY(lfkx_halo_f64(c, c->s)); }
	LFK_TRY(lfks_apply_a(c, dt, c->s, c->z));
	LFK_TRY(lfks_compact(c, c->z, (double*)c->staging, nullptr, nullptr));
	LFK_CUDA(c, cudaMemcpyAsync(out, c->staging, (size_t)nf * 8, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_apply_pressure(lfk_ctx *c, double dt) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfks_apply_pressure(c, dt);
}
extern "C" int lfk_correct(lfk_ctx *c, double dt) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_correct(c, dt);
}
extern "C" int lfk_extrapolate(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfks_extrapolate(c);
}
extern "C" int lfk_g2p(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_g2p(c);
}
extern "C" int lfk_cfl(lfk_ctx *c, double *value) {
	if (!c || !value) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_cfl(c, value);
}

// ---- fluid sources ------------------------------------------------------------------------------------------
static void free_sources(lfk_ctx *c) {
	dev_free(c->src_cell);
	dev_free(c->src_gcell);
	dev_free(c->src_of);
	dev_free(c->src_need);
	dev_free(c->src_vel);
	dev_free(c->src_target);
	c->src_entries = c->src_count = 0;
	c->src_active = c->src_coerce = false;
}

extern "C" int lfk_set_sources(lfk_ctx *c, const lfk_source *sources, uint64_t n) {
	if (!c || (!sources && n)) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, n < 65535, LFK_E_INVALID, "too many sources");
	const GridDesc &G = c->g;
	free_sources(c);
	struct Entry { uint32_t cell, gcell, src; };
	std::vector<Entry> ent;
	std::vector<double> vel(3 * (size_t)n);
	std::vector<uint32_t> target((size_t)n);
	std::vector<uint16_t> map;
	bool any = false, coerce = false;
	for (uint64_t k = 0; k < n; ++k) {
		const lfk_source &S = sources[k];
		LFK_REQUIRE(c, S.cells != nullptr || S.num_cells == 0, LFK_E_INVALID, "source without cells");
		for (int d = 0; d < 3; ++d) { vel[3 * k + d] = S.velocity[d]; }
		const uint64_t t = (uint64_t)S.target_density_cubic_root;
		target[k] = (uint32_t)(t * t * t);
		if (!S.active) { continue; }
		any = true;
		for (uint64_t j = 0; j < S.num_cells; ++j) {
			const uint64_t x = S.cells[3 * j], y = S.cells[3 * j + 1], z = S.cells[3 * j + 2];
			LFK_REQUIRE(c, x < (uint64_t)G.nx && y < (uint64_t)G.ny && z < (uint64_t)G.nz, LFK_E_INVALID,
				"source cell outside the grid");
			if ((long long)z < G.z0 || (long long)z >= G.z0 + G.nzl) { continue; } // another rank's slab
			const uint32_t local = (uint32_t)(x + (uint64_t)G.nx * (y + (uint64_t)G.ny * (z - (uint64_t)G.z0 + 1)));
			ent.push_back({ local, (uint32_t)(x + (uint64_t)G.nx * (y + (uint64_t)G.ny * z)), (uint32_t)k });
			if (S.coerce_velocity) {
				coerce = true;
				if (map.empty()) { map.assign((size_t)G.ncl, 0); }
				map[local] = (uint16_t)(k + 1); // the reference applies the sources in order: the last one wins
			}
		}
	}
	c->src_count = (uint32_t)n;
	c->src_active = any;
	// every rank must take the same branch in lfk_time_step, whether or not its own slab holds source cells
	for (uint64_t k = 0; k < n; ++k) { coerce |= sources[k].active && sources[k].coerce_velocity; }
	c->src_coerce = coerce;
	if (!any) { return 0; }
	std::stable_sort(ent.begin(), ent.end(), [](const Entry &a, const Entry &b) { return a.cell < b.cell; });
	const size_t ne = ent.size();
	c->src_entries = (uint32_t)ne;
	std::vector<uint32_t> cell(ne), gcell(ne), of(ne);
	for (size_t e = 0; e < ne; ++e) { cell[e] = ent[e].cell; gcell[e] = ent[e].gcell; of[e] = ent[e].src; }
	LFK_TRY(dev_alloc(c, &c->src_cell, ne));
	LFK_TRY(dev_alloc(c, &c->src_gcell, ne));
	LFK_TRY(dev_alloc(c, &c->src_of, ne));
	LFK_TRY(dev_alloc(c, &c->src_need, 2 * (ne + 1)));
	LFK_TRY(dev_alloc(c, &c->src_vel, 3 * (size_t)n));
	LFK_TRY(dev_alloc(c, &c->src_target, (size_t)n));
	if (ne) {
		LFK_CUDA(c, cudaMemcpyAsync(c->src_cell, cell.data(), ne * 4, cudaMemcpyHostToDevice, c->stream));
		LFK_CUDA(c, cudaMemcpyAsync(c->src_gcell, gcell.data(), ne * 4, cudaMemcpyHostToDevice, c->stream));
		LFK_CUDA(c, cudaMemcpyAsync(c->src_of, of.data(), ne * 4, cudaMemcpyHostToDevice, c->stream));
	}
	LFK_CUDA(c, cudaMemcpyAsync(c->src_vel, vel.data(), vel.size() * 8, cudaMemcpyHostToDevice, c->stream));
	LFK_CUDA(c, cudaMemcpyAsync(c->src_target, target.data(), target.size() * 4, cudaMemcpyHostToDevice, c->stream));
	if (coerce) {
		if (!c->src_map) { LFK_TRY(dev_alloc(c, &c->src_map, (size_t)G.ncl)); }
		if (map.empty()) { map.assign((size_t)G.ncl, 0); }
		LFK_CUDA(c, cudaMemcpyAsync(c->src_map, map.data(), map.size() * 2, cudaMemcpyHostToDevice, c->stream));
	}
	LFK_CUDA(c, cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
	return 0;
}
extern "C" int lfk_set_rng_seed(lfk_ctx *c, uint64_t seed) {
	if (!c) { return LFK_E_INVALID; }
	c->rng_seed = seed;
	c->rng_step = 0;
	return 0;
}
extern "C" int lfk_coerce_sources(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_coerce_sources(c);
}
extern "C" int lfk_update_sources(lfk_ctx *c, uint64_t *added) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_update_sources(c, added);
}

// ---- fused step: simulation::time_step(dt) (src/simulation.cpp:43-125) ---------------------------------------
// The reference sorts three times per step (:49, :62, :64); only the last sort feeds anything when there are no
// sources, so one sort after advection + collision is equivalent.  old_position never leaves registers.
extern "C" int lfk_time_step(lfk_ctx *c, double dt) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	if (c->src_active && c->src_coerce) { LFK_TRY(lfkp_coerce_sources(c)); } // :49 + :227-238
	LFK_TRY(lfkp_advect_collide(c, dt));            // :50-60
	LFK_TRY(lfkp_hash(c, c->tune.lean_sort != 0));  // :62-64 (lean: v / c are read through the permutation)
	if (c->src_active) {                            // :63-64: seed the source cells, sort again if anything was added
		uint64_t added = 0;
		LFK_TRY(lfkp_update_sources(c, &added));
		if (c->nranks > 1) { // every rank sorts again if ANY rank seeded (the sort's exchange is collective)
			double flag = added ? 1.0 : 0.0;
			LFK_CUDA(c, cudaMemcpyAsync(c->d_reduce, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
			LFK_TRY(lfkx_allreduce_max(c, c->d_reduce, 1));
			LFK_CUDA(c, cudaMemcpyAsync(&flag, c->d_reduce, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
			LFK_CUDA(c, cudaStreamSynchronize(c->stream));
			added = flag != 0.0 ? 1 : 0;
		}
		if (added) { LFK_TRY(lfkp_hash(c, c->tune.lean_sort != 0)); }
	}
	LFK_TRY(lfkg_p2g(c, dt, true));                 // :66-78 (gravity fused)
	LFK_TRY(lfks_solve(c, dt, nullptr, nullptr, true)); // :83-99 (initial guess: the previous step's pressure)
	LFK_TRY(lfks_apply_pressure(c, dt));            // :104
	LFK_TRY(lfkp_correct_collide(c, dt));           // :110-117
	LFK_TRY(lfks_extrapolate(c));                   // :119
	LFK_TRY(lfkp_g2p(c));                           // :121
	return 0;
}

extern "C" int lfk_time_step_cfl(lfk_ctx *c, double *dt_used) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	double cfl = 0.0;
	LFK_TRY(lfkp_cfl(c, &cfl));
	double dt = std::min(c->prm.cfl_number * cfl, 0.033); // src/simulation.cpp:127-129
	if (dt_used) { *dt_used = dt; }
	return lfk_time_step(c, dt);
}

extern "C" int lfk_update(lfk_ctx *c, double dt, uint64_t *substeps) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	uint64_t n = 0;
	while (true) { // src/simulation.cpp:31-41
		double cfl = 0.0;
		LFK_TRY(lfkp_cfl(c, &cfl));
		double ts = c->prm.cfl_number * cfl;
		++n;
		if (ts > dt) {
			LFK_TRY(lfk_time_step(c, dt));
			break;
		}
		LFK_TRY(lfk_time_step(c, ts));
		dt -= ts;
	}
	if (substeps) { *substeps = n; }
	return 0;
}

extern "C" int lfk_seed_box_device(lfk_ctx *c, const double start[3], const double size[3], const double velocity[3],
	uint32_t density, uint64_t seed, int append) {
	if (!c || !start || !size || !velocity || density < 1) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfkp_seed_box(c, start, size, velocity, density, seed, append);
}

extern "C" int lfk_synthetic_projection_device(lfk_ctx *c, uint64_t seed) {
	if (!c) { return LFK_E_INVALID; }
	NEED_PARAMS(c);
	return lfks_synthetic_projection(c, seed);
}

extern "C" int lfk_set_timing(lfk_ctx *c, int enabled) {
	if (!c) { return LFK_E_INVALID; }
	c->timing = enabled != 0;
	return 0;
}
extern "C" int lfk_set_tuning(lfk_ctx *c, const char *key, int value) {
	if (!c || !key) { return LFK_E_INVALID; }
	const std::string k(key);
	if (k == "p2g") { c->tune.p2g = value; }
	else if (k == "p2g_chunk") { c->tune.p2g_chunk = value; }
	else if (k == "mg_agg") { c->tune.mg_agg = value; }
	else if (k == "lean_sort") { c->tune.lean_sort = value; }
	else if (k == "mg_agg_cells") { c->tune.mg_agg_cells = value; }
	else if (k == "p2p") { c->tune.p2p = value; }
	else if (k == "ll_kb") { c->tune.ll_kb = value; c->pcg_graph_key = 0; }
	else if (k == "graph") { c->tune.graph = value; c->pcg_graph_key = 0; }
	else if (k == "mg_coarse") { c->tune.mg_coarse = value; c->pcg_graph_key = 0; }
	else if (k == "warm_start") { c->tune.warm_start = value; }
	else if (k == "red_blocks") { c->tune.red_blocks = value; }
	else { return lfk_fail(c, LFK_E_INVALID, "lfk_set_tuning: unknown key", __FILE__, __LINE__); }
	return 0;
}
extern "C" int lfk_get_stats(lfk_ctx *c, lfk_stats *out) {
	if (!c || !out) { return LFK_E_INVALID; }
	c->stats.num_particles = c->np;
	*out = c->stats;
	return 0;
}
extern "C" int lfk_reset_stats(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	memset(&c->stats, 0, sizeof(c->stats));
	return 0;
}
