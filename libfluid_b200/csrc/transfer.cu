// N3: host-mirror streaming and checkpoints.
//
// Callers of the reference read simulation::particles() every frame (testbed/main.cpp:52, the Maya node's particle
// cache, plugins/maya/nodes/grid_node.cpp:350-366) and save point clouds as text (include/fluid/data_structures/
// point_cloud.h:14-37).  With the state resident in HBM those become transfers, and this file is where they are made
// cheap:
//   * full particle records (152-byte AoS, the reference's layout) move in chunks through two staging slots on a
//     second stream, so the PCIe copy of chunk k overlaps the AoS <-> SoA kernel of chunk k +- 1 and the staging
//     memory is two chunks instead of a second copy of the whole particle set;
//   * positions only (what a renderer / mesher consumes: 24 of the 152 bytes) can be downloaded ASYNCHRONOUSLY into
//     pinned memory: the copy runs on the second stream while the next time step computes;
//   * a binary checkpoint (SoA fields of the particles, the slab's cells, the solver's warm-start state) replaces the
//     text point cloud for restarts: restoring it reproduces the continued run bit for bit.
#include "lfk_internal.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

// particles per chunk: 8 M (1.2 GB of AoS records); LFK_XFER_CHUNK overrides it (tests exercise the multi-chunk path)
static uint64_t xfer_chunk() {
	if (const char *env = getenv("LFK_XFER_CHUNK")) {
		const long long v = atoll(env);
		if (v > 0) { return (uint64_t)v; }
	}
	return 8ull << 20;
}

static int xfer_init(lfk_ctx *c) {
	if (c->copy_stream) { return 0; }
	LFK_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int k = 0; k < 6; ++k) { LFK_CUDA(c, cudaEventCreateWithFlags(&c->xfer_ev[k], cudaEventDisableTiming)); }
	return 0;
}

int lfkt_destroy(lfk_ctx *c) {
	if (c->copy_stream) {
		cudaStreamSynchronize(c->copy_stream);
		for (int k = 0; k < 6; ++k) {
			if (c->xfer_ev[k]) { cudaEventDestroy(c->xfer_ev[k]); c->xfer_ev[k] = nullptr; }
		}
		cudaStreamDestroy(c->copy_stream);
		c->copy_stream = nullptr;
	}
	if (c->pos_stage) { cudaFree(c->pos_stage); c->pos_stage = nullptr; c->pos_stage_bytes = 0; }
	return 0;
}

// host AoS -> device SoA: copy stream: H2D of chunk k into slot k & 1 | main stream: AoS -> SoA kernel of chunk k
int lfkt_upload_particles_pipelined(lfk_ctx *c, const void *aos152, uint64_t n) {
	LFK_TRY(xfer_init(c));
	const uint64_t chunk = n < xfer_chunk() ? n : xfer_chunk();
	LFK_TRY(lfk_reserve_staging(c, (size_t)(n <= chunk ? chunk : 2 * chunk) * 152));
	cudaEvent_t *filled = c->xfer_ev, *freed = c->xfer_ev + 2;
	// the staging buffer may still be in use by earlier work of the main stream
	LFK_CUDA(c, cudaEventRecord(freed[0], c->stream));
	LFK_CUDA(c, cudaEventRecord(freed[1], c->stream));
	uint64_t k = 0;
	for (uint64_t at = 0; at < n; at += chunk, ++k) {
		const uint64_t m = n - at < chunk ? n - at : chunk;
		char *slot = (char*)c->staging + (size_t)(k & 1) * chunk * 152;
		LFK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, freed[k & 1], 0));
		LFK_CUDA(c, cudaMemcpyAsync(slot, (const char*)aos152 + (size_t)at * 152, (size_t)m * 152, cudaMemcpyHostToDevice,
			c->copy_stream));
		LFK_CUDA(c, cudaEventRecord(filled[k & 1], c->copy_stream));
		LFK_CUDA(c, cudaStreamWaitEvent(c->stream, filled[k & 1], 0));
		LFK_TRY(lfkp_aos_to_soa(c, slot, m, at));
		LFK_CUDA(c, cudaEventRecord(freed[k & 1], c->stream));
	}
	// host pointers are borrowed for the call only: the copies must have left the caller's buffer
	LFK_CUDA(c, cudaStreamSynchronize(c->copy_stream));
	return 0;
}

// device SoA -> host AoS: main stream: SoA -> AoS kernel of chunk k into slot k & 1 | copy stream: D2H of chunk k
int lfkt_download_particles_pipelined(lfk_ctx *c, void *aos152, uint64_t n) {
	LFK_TRY(xfer_init(c));
	const uint64_t chunk = n < xfer_chunk() ? n : xfer_chunk();
	LFK_TRY(lfk_reserve_staging(c, (size_t)(n <= chunk ? chunk : 2 * chunk) * 152));
	cudaEvent_t *filled = c->xfer_ev, *freed = c->xfer_ev + 2;
	LFK_CUDA(c, cudaEventRecord(freed[0], c->copy_stream));
	LFK_CUDA(c, cudaEventRecord(freed[1], c->copy_stream));
	uint64_t k = 0;
	for (uint64_t at = 0; at < n; at += chunk, ++k) {
		const uint64_t m = n - at < chunk ? n - at : chunk;
		char *slot = (char*)c->staging + (size_t)(k & 1) * chunk * 152;
		LFK_CUDA(c, cudaStreamWaitEvent(c->stream, freed[k & 1], 0));
		LFK_TRY(lfkp_soa_to_aos(c, slot, m, at));
		LFK_CUDA(c, cudaEventRecord(filled[k & 1], c->stream));
		LFK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, filled[k & 1], 0));
		LFK_CUDA(c, cudaMemcpyAsync((char*)aos152 + (size_t)at * 152, slot, (size_t)m * 152, cudaMemcpyDeviceToHost,
			c->copy_stream));
		LFK_CUDA(c, cudaEventRecord(freed[k & 1], c->copy_stream));
	}
	LFK_CUDA(c, cudaStreamSynchronize(c->copy_stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

// ---- asynchronous positions download ----------------------------------------------------------------------------
extern "C" int lfk_host_alloc(void **out, uint64_t bytes) {
	if (!out) { return LFK_E_INVALID; }
	*out = nullptr;
	cudaError_t e = cudaHostAlloc(out, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
	if (e != cudaSuccess) { return lfk_fail(nullptr, -(int)e, cudaGetErrorString(e), __FILE__, __LINE__); }
	return 0;
}
extern "C" int lfk_host_free(void *p) {
	if (p) { cudaFreeHost(p); }
	return 0;
}

extern "C" int lfk_download_positions_async(lfk_ctx *c, double *xyz, uint64_t capacity, uint64_t *n) {
	if (!c) { return LFK_E_INVALID; }
	if (n) { *n = c->np; }
	LFK_REQUIRE(c, capacity >= c->np, LFK_E_CAPACITY, "position buffer too small");
	if (c->np == 0) { return 0; }
	LFK_REQUIRE(c, xyz != nullptr, LFK_E_INVALID, "NULL position buffer");
	LFK_TRY(xfer_init(c));
	cudaEvent_t ready = c->xfer_ev[4], copied = c->xfer_ev[5];
	if (c->pos_pending) { // the previous download still owns the device buffer
		LFK_CUDA(c, cudaStreamWaitEvent(c->stream, copied, 0));
	}
	const size_t bytes = (size_t)c->np * 24;
	if (bytes > c->pos_stage_bytes) {
		if (c->pos_pending) { LFK_CUDA(c, cudaEventSynchronize(copied)); }
		if (c->pos_stage) { cudaFree(c->pos_stage); c->pos_stage = nullptr; c->pos_stage_bytes = 0; }
		const size_t want = bytes + bytes / 16;
		LFK_CUDA(c, cudaMalloc(&c->pos_stage, want));
		c->pos_stage_bytes = want;
	}
	LFK_TRY(lfkp_positions_to_aos(c, (double*)c->pos_stage, c->np));
	LFK_CUDA(c, cudaEventRecord(ready, c->stream));
	LFK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, ready, 0));
	// In pieces: the step that runs meanwhile reads a handful of scalars back per solve (convergence flag, CFL maximum),
	// and a device-to-host engine serves its queue in order -- behind one 3 GB copy those reads stall the whole step for
	// the duration of the copy (measured: 127 ms per step instead of 65, tools/stream_probe.py); behind a 4 MB piece
	// they wait 80 us.
	const size_t piece = 4u << 20;
	for (size_t at = 0; at < bytes; at += piece) {
		const size_t m = bytes - at < piece ? bytes - at : piece;
		LFK_CUDA(c, cudaMemcpyAsync((char*)xyz + at, (const char*)c->pos_stage + at, m, cudaMemcpyDeviceToHost, c->copy_stream));
	}
	LFK_CUDA(c, cudaEventRecord(copied, c->copy_stream));
	c->pos_pending = true;
	return 0;
}

extern "C" int lfk_wait_transfers(lfk_ctx *c) {
	if (!c) { return LFK_E_INVALID; }
	if (c->pos_pending) {
		LFK_CUDA(c, cudaEventSynchronize(c->xfer_ev[5]));
		c->pos_pending = false;
	}
	return 0;
}

// ---- checkpoints ------------------------------------------------------------------------------------------------
namespace {
struct CkptHeader {
	char magic[8];           // "LFKCKPT1"
	uint64_t nx, ny, nz;     // whole grid
	uint64_t z0, nzl;        // the slab this file holds
	uint64_t nranks, rank;
	uint64_t np;             // particles of this rank
	uint64_t has_old_grid, has_pressure;
	double last_solve_dt;
	uint64_t last_iters, last_solve_ok;
	uint64_t rng_seed, rng_step;
	lfk_params params;
};

// device -> file / file -> device through a pinned bounce buffer
struct Bounce {
	lfk_ctx *c;
	void *host = nullptr;
	size_t bytes = 64u << 20;
	explicit Bounce(lfk_ctx *ctx) : c(ctx) { if (cudaHostAlloc(&host, bytes, cudaHostAllocDefault) != cudaSuccess) { host = nullptr; cudaGetLastError(); } }
	~Bounce() { if (host) { cudaFreeHost(host); } }
	int write(FILE *f, const void *dev, size_t n) {
		for (size_t at = 0; at < n; at += bytes) {
			const size_t m = n - at < bytes ? n - at : bytes;
			LFK_CUDA(c, cudaMemcpyAsync(host, (const char*)dev + at, m, cudaMemcpyDeviceToHost, c->stream));
			LFK_CUDA(c, cudaStreamSynchronize(c->stream));
			LFK_REQUIRE(c, fwrite(host, 1, m, f) == m, LFK_E_INVALID, "checkpoint: short write");
		}
		return 0;
	}
	int read(FILE *f, void *dev, size_t n) {
		for (size_t at = 0; at < n; at += bytes) {
			const size_t m = n - at < bytes ? n - at : bytes;
			LFK_REQUIRE(c, fread(host, 1, m, f) == m, LFK_E_INVALID, "checkpoint: short read");
			LFK_CUDA(c, cudaMemcpyAsync((char*)dev + at, host, m, cudaMemcpyHostToDevice, c->stream));
			LFK_CUDA(c, cudaStreamSynchronize(c->stream));
		}
		return 0;
	}
};

std::string ckpt_path(const lfk_ctx *c, const char *path) {
	std::string p(path);
	if (c->nranks > 1) { p += ".rank" + std::to_string(c->rank); }
	return p;
}
}

extern "C" int lfk_checkpoint_save(lfk_ctx *c, const char *path) {
	if (!c || !path) { return LFK_E_INVALID; }
	LFK_TRY(lfkp_materialise_vc(c)); // every field in particle order
	FILE *f = fopen(ckpt_path(c, path).c_str(), "wb");
	LFK_REQUIRE(c, f != nullptr, LFK_E_INVALID, "checkpoint: cannot open the file for writing");
	Bounce B(c);
	int rc = 0;
	do {
		if (!B.host) { rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: no pinned memory", __FILE__, __LINE__); break; }
		CkptHeader H{};
		memcpy(H.magic, "LFKCKPT1", 8);
		H.nx = (uint64_t)c->g.nx; H.ny = (uint64_t)c->g.ny; H.nz = (uint64_t)c->g.nz;
		H.z0 = (uint64_t)c->g.z0; H.nzl = (uint64_t)c->g.nzl;
		H.nranks = (uint64_t)c->nranks; H.rank = (uint64_t)c->rank;
		H.np = c->np;
		H.has_old_grid = c->vel_old[0] ? 1 : 0;
		H.has_pressure = c->pressure_valid ? 1 : 0;
		H.last_solve_dt = c->last_solve_dt; H.last_iters = c->last_iters; H.last_solve_ok = c->last_solve_ok ? 1 : 0;
		H.rng_seed = c->rng_seed; H.rng_step = c->rng_step;
		H.params = c->prm;
		if (fwrite(&H, sizeof(H), 1, f) != 1) { rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: short write", __FILE__, __LINE__); break; }
		// particles: position, velocity, c rows (old_position == position between steps), own particles only
		for (int fld = 0; fld < 15 && rc == 0; ++fld) { rc = B.write(f, c->P.f[fld] + c->first, (size_t)c->np * 8); }
		if (rc) { break; }
		// cells of the slab incl. its ghost layers (they are part of the device state between steps)
		const size_t ncl = (size_t)c->g.ncl;
		for (int d = 0; d < 3 && rc == 0; ++d) { rc = B.write(f, c->vel[d], ncl * 8); }
		if (rc == 0) { rc = B.write(f, c->typ, ncl); }
		if (H.has_old_grid) {
			for (int d = 0; d < 3 && rc == 0; ++d) { rc = B.write(f, c->vel_old[d], ncl * 8); }
		}
		if (H.has_pressure && rc == 0) { rc = B.write(f, c->p, ncl * 8); }
	} while (0);
	if (fclose(f) != 0 && rc == 0) { rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: close failed", __FILE__, __LINE__); }
	return rc;
}

extern "C" int lfk_checkpoint_load(lfk_ctx *c, const char *path) {
	if (!c || !path) { return LFK_E_INVALID; }
	c->speed2_valid = false;
	FILE *f = fopen(ckpt_path(c, path).c_str(), "rb");
	LFK_REQUIRE(c, f != nullptr, LFK_E_INVALID, "checkpoint: cannot open the file");
	Bounce B(c);
	int rc = 0;
	do {
		if (!B.host) { rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: no pinned memory", __FILE__, __LINE__); break; }
		CkptHeader H{};
		if (fread(&H, sizeof(H), 1, f) != 1 || memcmp(H.magic, "LFKCKPT1", 8) != 0) {
			rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: not an lfk checkpoint", __FILE__, __LINE__);
			break;
		}
		if (H.nx != (uint64_t)c->g.nx || H.ny != (uint64_t)c->g.ny || H.nz != (uint64_t)c->g.nz ||
			H.nranks != (uint64_t)c->nranks || H.rank != (uint64_t)c->rank) {
			rc = lfk_fail(c, LFK_E_INVALID, "checkpoint: grid or rank layout differs from this context", __FILE__, __LINE__);
			break;
		}
		if ((rc = lfk_set_params(c, &H.params)) != 0) { break; }
		c->np = 0; c->first = 0; c->ntot = 0;
		c->v_deferred = c->c_deferred = false;
		if ((rc = lfkp_reserve_particles(c, H.np)) != 0) { break; }
		for (int fld = 0; fld < 15 && rc == 0; ++fld) { rc = B.read(f, c->P.f[fld], (size_t)H.np * 8); }
		if (rc) { break; }
		c->np = c->ntot = H.np;
		c->old_valid = false;
		c->table_valid = false;
		c->keys_valid = false;
		const size_t ncl = (size_t)c->g.ncl;
		for (int d = 0; d < 3 && rc == 0; ++d) { rc = B.read(f, c->vel[d], ncl * 8); }
		if (rc == 0) { rc = B.read(f, c->typ, ncl); }
		if (H.has_old_grid) {
			if (!c->vel_old[0]) { rc = lfk_fail(c, LFK_E_STATE, "checkpoint: FLIP snapshot without FLIP parameters", __FILE__, __LINE__); break; }
			for (int d = 0; d < 3 && rc == 0; ++d) { rc = B.read(f, c->vel_old[d], ncl * 8); }
		}
		if (H.has_pressure && rc == 0) { rc = B.read(f, c->p, ncl * 8); }
		c->pressure_valid = H.has_pressure != 0;
		c->system_valid = false;
		c->last_solve_dt = H.last_solve_dt; c->last_iters = H.last_iters; c->last_solve_ok = H.last_solve_ok != 0;
		c->rng_seed = H.rng_seed; c->rng_step = H.rng_step;
	} while (0);
	fclose(f);
	return rc;
}
