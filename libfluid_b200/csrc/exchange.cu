// Multi-GPU plumbing: one context per rank owns a z-slab; neighbours exchange one ghost layer of cell data and
// the PCG scalars are all-reduced.  NCCL is resolved at run time with dlopen (libnccl.so.2 -- inside a PyTorch
// process that is torch's own copy), so single-GPU use has no NCCL dependency at all.
#include "lfk_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
NcclApi g_nccl;

bool load_nccl() {
	if (g_nccl.ok) { return true; }
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char *n : names) {
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) { break; }
	}
	if (!g_nccl.handle) { return false; }
#define SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.handle, name); if (!g_nccl.field) { return false; }
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(AllReduce, "ncclAllReduce");
	SYM(AllGather, "ncclAllGather");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.ok = true;
	return true;
}
}

#define LFK_NCCL(ctx, expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) { \
	return lfk_fail((ctx), -(1000 + (int)r__), g_nccl.GetErrorString(r__), __FILE__, __LINE__); } } while (0)

extern "C" int lfk_nccl_unique_id(void *out128) {
	if (!out128) { return LFK_E_INVALID; }
	if (!load_nccl()) { return lfk_fail(nullptr, LFK_E_NCCL, "libnccl.so.2 could not be loaded", __FILE__, __LINE__); }
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r != ncclSuccess) { return lfk_fail(nullptr, -(1000 + (int)r), g_nccl.GetErrorString(r), __FILE__, __LINE__); }
	memcpy(out128, &id, 128);
	return 0;
}

// =========================================================================================================
// Halos through peer memory.
//
// A halo exchange between z neighbours moves one cell layer (at most nx * ny doubles) each way.  As an NCCL send / recv
// group it costs ~13 us of launch and handshake for a few microseconds of NVLink traffic, and one PCG iteration needs
// ~75 of them (every multigrid half-sweep on every level): 2.06 ms per iteration on two GPUs against 1.03 ms on one
// (profiles/r2a_bench_2gpu.json).  Here every rank owns an ARENA in its own HBM that its two neighbours map with CUDA IPC;
// one kernel per exchange
//   1. stores the rank's top / bottom owned layer straight into the upper / lower neighbour's arena (NVLink stores),
//   2. publishes them: system-scope fence, then -- by the last block to finish -- a release store of the exchange's
//      epoch number into the neighbour's flag word,
//   3. waits until its own two flag words have reached the epoch (acquire loads of local memory),
//   4. copies what the neighbours stored from the arena into the ghost layers.
// Slots alternate with the parity of the epoch.  That is enough to rule out overwriting a slot that is still being
// read: a neighbour can start exchange e + 1 only after finishing its own exchange e, which waited for this rank's
// flag of exchange e, which this rank raised only after finishing exchange e - 1 -- the last reader of the slot that
// e + 1 writes.  Every rank issues the same sequence of exchanges (as for NCCL), so the epochs agree by construction.
// A spin that lasts longer than ~4 s raises an error flag instead of hanging the device.
// =========================================================================================================
#define LL_MAX_BYTES (2048u * 1024u) // capacity of a flag-in-data slot; layers above lfk_tuning::ll_kb take the fence-and-flag protocol
#define ARENA_HEADER 256 // bytes: flag words + block counter + error word
#define ARENA_MAX_RANKS 64
struct ArenaHeader {
	unsigned long long sig[2]; // [0] raised by the lower neighbour, [1] by the upper one
	unsigned long long sepoch; // scalar all-reduces completed by this rank
	unsigned long long epoch;  // exchanges completed by this rank (device resident, so that the kernel's arguments never
	                           // change and a captured CUDA graph of a PCG iteration can be replayed)
	unsigned counter;          // blocks of the running exchange kernel that have published their stores
	unsigned error;            // != 0: a wait timed out
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
// grid-strided copy of `bytes` (a multiple of 4; 16-byte vectors when everything is aligned); SRC_ARENA: the source
// was written by another GPU -- bypass L1
template <bool SRC_ARENA> __device__ __forceinline__ void halo_copy(char *dst, const char *src, size_t bytes) {
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if (((((size_t)dst) | ((size_t)src) | bytes) & 15u) == 0) {
		const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
		uint4 *d4 = reinterpret_cast<uint4 *>(dst);
		for (size_t i = tid; i < bytes / 16; i += nth) { d4[i] = SRC_ARENA ? __ldcg(s4 + i) : s4[i]; }
	} else if (((((size_t)dst) | ((size_t)src) | bytes) & 3u) == 0) {
		const unsigned *s1 = reinterpret_cast<const unsigned *>(src);
		unsigned *d1 = reinterpret_cast<unsigned *>(dst);
		for (size_t i = tid; i < bytes / 4; i += nth) { d1[i] = SRC_ARENA ? __ldcg(s1 + i) : s1[i]; }
	} else {
		for (size_t i = tid; i < bytes; i += nth) { dst[i] = SRC_ARENA ? __ldcg(src + i) : src[i]; }
	}
}

// up / dn: the neighbours' arenas (NULL at the ends of the rank chain); top / bottom: this rank's boundary layers;
// ghost_top / ghost_bottom: where the neighbours' layers go
__global__ void __launch_bounds__(256) k_halo_p2p(char *mine, char *up, char *dn, const char *top, const char *bottom,
	char *ghost_top, char *ghost_bottom, size_t bytes, size_t slot) {
	ArenaHeader *H = reinterpret_cast<ArenaHeader *>(mine);
	// every block reads the epoch before it arrives at the counter below, and the last block to arrive is the one that
	// advances it: all blocks of a launch see the same number
	const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(&H->epoch) + 1ull;
	const size_t par = (size_t)(epoch & 1ull);
	// receive slots of an arena: [from lower][parity], [from upper][parity]
	if (up) { halo_copy<false>(up + ARENA_HEADER + (0 * 2 + par) * slot, top, bytes); }       // I am the upper rank's LOWER neighbour
	if (dn) { halo_copy<false>(dn + ARENA_HEADER + (1 * 2 + par) * slot, bottom, bytes); }    // ... the lower rank's UPPER neighbour
	__threadfence_system();
	__syncthreads();
	__shared__ bool last;
	if (threadIdx.x == 0) {
		const unsigned t = atomicAdd(&H->counter, 1u);
		last = t == gridDim.x - 1;
		if (last) {
			H->counter = 0; // every block has arrived; the next exchange is a later launch on the same stream
			*reinterpret_cast<volatile unsigned long long *>(&H->epoch) = epoch;
			__threadfence_system();
			if (up) { st_release_sys(&reinterpret_cast<ArenaHeader *>(up)->sig[0], epoch); }
			if (dn) { st_release_sys(&reinterpret_cast<ArenaHeader *>(dn)->sig[1], epoch); }
		}
		const long long t0 = clock64();
		bool ok = true;
		while ((dn && ld_acquire_sys(&H->sig[0]) < epoch) || (up && ld_acquire_sys(&H->sig[1]) < epoch)) {
			if (clock64() - t0 > 8000000000ll) { // ~4 s at 1.97 GHz: a neighbour is gone
				H->error = 1u;
				ok = false;
				break;
			}
		}
		(void)ok;
	}
	__syncthreads();
	if (dn) { halo_copy<true>(ghost_bottom, mine + ARENA_HEADER + (0 * 2 + par) * slot, bytes); }
	if (up) { halo_copy<true>(ghost_top, mine + ARENA_HEADER + (1 * 2 + par) * slot, bytes); }
}

// Small layers (the coarse multigrid levels: a few KB) are pure latency, so they use a flag-in-data protocol instead:
// every 4-byte element travels as an 8-byte word {element, epoch}; an aligned 8-byte store arrives whole, so the receiver
// simply polls each word until its epoch matches -- no fence, no separate flag, one NVLink traversal per exchange.
__device__ __forceinline__ void st_ll(char *slot, size_t i, unsigned v, unsigned flag) {
	asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot + 8 * i), "r"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ld_ll(const char *slot, size_t i, unsigned flag, unsigned &v, unsigned *error) {
	unsigned f;
	long long t0 = 0;
	for (unsigned spins = 0;; ++spins) {
		asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(slot + 8 * i) : "memory");
		if (f == flag) { return true; }
		if ((spins & 1023u) == 1023u) {
			if (t0 == 0) { t0 = clock64(); }
			else if (clock64() - t0 > 8000000000ll) { *error = 1u; return false; } // ~4 s: a neighbour is gone
		}
	}
}
__global__ void __launch_bounds__(256) k_halo_ll(char *mine, char *up, char *dn, const unsigned *top,
	const unsigned *bottom, unsigned *ghost_top, unsigned *ghost_bottom, size_t nwords, size_t ll_base) {
	ArenaHeader *H = reinterpret_cast<ArenaHeader *>(mine);
	const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(&H->epoch) + 1ull;
	const size_t par = (size_t)(epoch & 1ull);
	const unsigned flag = (unsigned)epoch; // (arena words start at 0 and epochs at 1)
	const size_t slot = 2 * LL_MAX_BYTES;
	mine += ll_base;
	if (up) { up += ll_base; }
	if (dn) { dn += ll_base; }
	H = reinterpret_cast<ArenaHeader *>(mine - ll_base);
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
	if (up) {
		char *dst = up + ARENA_HEADER + (0 * 2 + par) * slot;
		for (size_t i = tid; i < nwords; i += nth) { st_ll(dst, i, top[i], flag); }
	}
	if (dn) {
		char *dst = dn + ARENA_HEADER + (1 * 2 + par) * slot;
		for (size_t i = tid; i < nwords; i += nth) { st_ll(dst, i, bottom[i], flag); }
	}
	if (dn) {
		const char *src = mine + ARENA_HEADER + (0 * 2 + par) * slot;
		for (size_t i = tid; i < nwords; i += nth) {
			unsigned v;
			if (!ld_ll(src, i, flag, v, &H->error)) { break; }
			ghost_bottom[i] = v;
		}
	}
	if (up) {
		const char *src = mine + ARENA_HEADER + (1 * 2 + par) * slot;
		for (size_t i = tid; i < nwords; i += nth) {
			unsigned v;
			if (!ld_ll(src, i, flag, v, &H->error)) { break; }
			ghost_top[i] = v;
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) { // the last block to finish advances the epoch (see k_halo_p2p)
		const unsigned t = atomicAdd(&H->counter, 1u);
		if (t == gridDim.x - 1) {
			H->counter = 0;
			*reinterpret_cast<volatile unsigned long long *>(&H->epoch) = epoch;
		}
	}
}

static void arena_unmap(lfk_ctx *c) {
	for (size_t r = 0; r < c->arena_all.size(); ++r) {
		if ((int)r != c->rank && c->arena_all[r]) { cudaIpcCloseMemHandle(c->arena_all[r]); }
	}
	c->arena_all.clear();
	c->arena_peer[0] = c->arena_peer[1] = nullptr;
	if (c->arena_all_d) { cudaFree(c->arena_all_d); c->arena_all_d = nullptr; }
}

// ---- scalar all-reduce over peer memory, fused with the PCG finaliser ----------------------------------------------
// Every rank stores its partial value (as two {half, epoch} words) into EVERY rank's arena, then folds the nranks
// values it received in rank order -- every rank computes the same sum bit for bit -- and runs the finaliser that the
// single-GPU kernels run in their last block.  One launch of one warp-sized block instead of ncclAllReduce + k_finalize.
__global__ void __launch_bounds__(ARENA_MAX_RANKS) k_allreduce_ll(char *mine, char *const *all, int nranks, int rank,
	size_t base, double *field, int is_max, PcgScalars *scal, int which) {
	ArenaHeader *H = reinterpret_cast<ArenaHeader *>(mine);
	__shared__ double vals[ARENA_MAX_RANKS];
	const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(&H->sepoch) + 1ull;
	const size_t par = (size_t)(epoch & 1ull);
	const unsigned flag = (unsigned)epoch;
	const int t = (int)threadIdx.x;
	if (t < nranks) {
		const unsigned long long bits = (unsigned long long)__double_as_longlong(*field);
		char *dst = all[t] + base + (par * ARENA_MAX_RANKS + (size_t)rank) * 16;
		st_ll(dst, 0, (unsigned)bits, flag);
		st_ll(dst, 1, (unsigned)(bits >> 32), flag);
		const char *src = mine + base + (par * ARENA_MAX_RANKS + (size_t)t) * 16;
		unsigned lo = 0, hi = 0;
		const bool ok = ld_ll(src, 0, flag, lo, &H->error) && ld_ll(src, 1, flag, hi, &H->error);
		vals[t] = ok ? __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo)) : 0.0;
	}
	__syncthreads();
	if (t == 0) {
		double acc = vals[0];
		for (int r = 1; r < nranks; ++r) { acc = is_max ? fmax(acc, vals[r]) : acc + vals[r]; }
		*field = acc;
		*reinterpret_cast<volatile unsigned long long *>(&H->sepoch) = epoch;
		if (which == FIN_BB || !scal->done) { pcg_finalize(scal, which); }
	}
}

int lfkx_allreduce_finalize(lfk_ctx *c, double *field, bool is_max, int which) {
	if (!(c->p2p && c->tune.p2p && c->arena_all_d)) { return 0; }
	const size_t base = ARENA_HEADER + 4 * c->arena_slot + 4 * (size_t)(2 * LL_MAX_BYTES);
	k_allreduce_ll<<<1, ARENA_MAX_RANKS, 0, c->stream>>>(c->arena, c->arena_all_d, c->nranks, c->rank, base, field,
		is_max ? 1 : 0, c->d_scal, which);
	++c->stats.kernel_launches;
	if (cudaGetLastError() != cudaSuccess) { return 0; }
	return 1;
}

static int arena_setup(lfk_ctx *c) {
	c->p2p = false;
	if (c->nranks == 1) { return 0; }
	if (const char *env = getenv("LFK_P2P")) {
		if (atoi(env) == 0) { return 0; }
	}
	ncclComm_t comm = (ncclComm_t)c->comm;
	c->arena_slot = (((size_t)c->g.sxy * sizeof(double)) + 255) / 256 * 256;
	// [header | 2 x 2 fence-protocol slots of one fp64 layer | 2 x 2 flag-in-data slots] -- the two protocols never share
	// a slot, so a raw data word can never be mistaken for an {element, epoch} word
	// ... | scalar all-reduce words [parity][rank][2]]
	const size_t bytes = ARENA_HEADER + 4 * c->arena_slot + 4 * (size_t)(2 * LL_MAX_BYTES) + 2 * ARENA_MAX_RANKS * 16;
	LFK_CUDA(c, cudaMalloc((void**)&c->arena, bytes));
	LFK_CUDA(c, cudaMemsetAsync(c->arena, 0, bytes, c->stream));
	// exchange the IPC handles of all arenas (64 bytes each) with an NCCL all-gather; `ok` words tell every rank whether
	// EVERY rank could map its neighbours, so that all of them take the same path
	cudaIpcMemHandle_t mine;
	bool have = cudaIpcGetMemHandle(&mine, c->arena) == cudaSuccess;
	if (!have) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); }
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	char *d_all = nullptr;
	LFK_CUDA(c, cudaMalloc((void**)&d_all, (size_t)c->nranks * 64 + 64));
	LFK_CUDA(c, cudaMemcpyAsync(d_all + (size_t)c->nranks * 64, &mine, 64, cudaMemcpyHostToDevice, c->stream));
	LFK_NCCL(c, g_nccl.AllGather(d_all + (size_t)c->nranks * 64, d_all, 64, ncclChar, comm, c->stream));
	std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
	LFK_CUDA(c, cudaMemcpyAsync(all.data(), d_all, (size_t)c->nranks * 64, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	bool ok = have && c->nranks <= ARENA_MAX_RANKS;
	c->arena_all.assign((size_t)c->nranks, nullptr);
	c->arena_all[(size_t)c->rank] = c->arena;
	for (int r = 0; r < c->nranks && ok; ++r) { // every rank maps every arena (the scalar all-reduce writes to all of them)
		if (r == c->rank) { continue; }
		void *ptr = nullptr;
		if (cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
			cudaGetLastError();
			ok = false;
			break;
		}
		c->arena_all[(size_t)r] = (char*)ptr;
	}
	if (ok) {
		if (c->rank + 1 < c->nranks) { c->arena_peer[0] = c->arena_all[(size_t)c->rank + 1]; }
		if (c->rank > 0) { c->arena_peer[1] = c->arena_all[(size_t)c->rank - 1]; }
		LFK_CUDA(c, cudaMalloc((void**)&c->arena_all_d, (size_t)c->nranks * sizeof(char*)));
		LFK_CUDA(c, cudaMemcpyAsync(c->arena_all_d, c->arena_all.data(), (size_t)c->nranks * sizeof(char*),
			cudaMemcpyHostToDevice, c->stream));
	}
	// agree: sum of (ok ? 0 : 1) over the ranks must be 0
	double flag = ok ? 0.0 : 1.0;
	double *d_flag = (double*)d_all; // reuse
	LFK_CUDA(c, cudaMemcpyAsync(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	LFK_NCCL(c, g_nccl.AllReduce(d_flag, d_flag, 1, ncclDouble, ncclSum, comm, c->stream));
	LFK_CUDA(c, cudaMemcpyAsync(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream)); // (also: every arena is zeroed before anybody can write into it)
	cudaFree(d_all);
	c->p2p = flag == 0.0;
	if (!c->p2p) { arena_unmap(c); }
	c->halo_epoch = 0;
	return 0;
}

int lfkx_init(lfk_ctx *c, const void *nccl_id128) {
	if (c->nranks == 1) { return 0; }
	LFK_REQUIRE(c, nccl_id128 != nullptr, LFK_E_INVALID, "nranks > 1 needs an NCCL unique id");
	LFK_REQUIRE(c, load_nccl(), LFK_E_NCCL, "libnccl.so.2 could not be loaded");
	ncclUniqueId id;
	memcpy(&id, nccl_id128, 128);
	ncclComm_t comm;
	LFK_NCCL(c, g_nccl.CommInitRank(&comm, c->nranks, id, c->rank));
	c->comm = comm;
	return arena_setup(c);
}

// rendezvous of all ranks on the context's stream (a one-element all-reduce + synchronisation)
static void comm_barrier(lfk_ctx *c) {
	if (!c->comm || !c->d_reduce) { return; }
	if (g_nccl.AllReduce(c->d_reduce, c->d_reduce, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream) == ncclSuccess) {
		cudaStreamSynchronize(c->stream);
	}
}

int lfkx_destroy(lfk_ctx *c) {
	if (c->p2p) {
		// CUDA IPC: an importer must close its mapping before the exporter frees the allocation.  Every rank first
		// reaches this point (nobody issues another exchange), then unmaps its peers, then -- after a second rendezvous
		// -- frees its own arena.
		comm_barrier(c);
		arena_unmap(c);
		comm_barrier(c);
	} else {
		arena_unmap(c);
	}
	if (c->comm) {
		g_nccl.CommDestroy((ncclComm_t)c->comm);
		c->comm = nullptr;
	}
	if (c->arena) { cudaFree(c->arena); c->arena = nullptr; }
	c->p2p = false;
	return 0;
}

// != 0 once an exchange timed out waiting for a neighbour
int lfkx_check(lfk_ctx *c) {
	if (!c->p2p) { return 0; }
	unsigned err = 0;
	LFK_CUDA(c, cudaMemcpyAsync(&err, c->arena + offsetof(ArenaHeader, error), sizeof(unsigned), cudaMemcpyDeviceToHost,
		c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	LFK_REQUIRE(c, err == 0, LFK_E_NCCL, "a peer-memory halo exchange timed out waiting for a neighbour rank");
	return 0;
}

// fills the two z ghost layers of a per-cell array: my top owned layer -> upper neighbour's bottom ghost, my bottom
// owned layer -> lower neighbour's top ghost
static int halo_bytes(lfk_ctx *c, void *field, size_t layer_elems, int nzl, ncclDataType_t dt, size_t esz) {
	if (c->nranks == 1) { return 0; }
	PhaseTimer T(c, LFK_PHASE_EXCHANGE);
	ncclComm_t comm = (ncclComm_t)c->comm;
	char *f = (char*)field;
	size_t L = layer_elems * esz;
	if (c->p2p && c->tune.p2p && L % 4 == 0 && L <= LL_MAX_BYTES && L <= (size_t)c->tune.ll_kb * 1024u &&
		(((size_t)f | (size_t)(f + L)) & 3u) == 0) {
		++c->halo_epoch;
		unsigned nb = (unsigned)((L / 4 + 255) / 256);
		nb = nb < 1 ? 1 : (nb > 128 ? 128 : nb);
		LFK_LAUNCH(c, k_halo_ll, nb, 256, 0, c->arena, c->arena_peer[0], c->arena_peer[1], (const unsigned*)(f + (size_t)nzl * L),
			(const unsigned*)(f + L), (unsigned*)(f + (size_t)(nzl + 1) * L), (unsigned*)f, L / 4, 4 * c->arena_slot);
		return 0;
	}
	if (c->p2p && c->tune.p2p && L <= c->arena_slot) {
		++c->halo_epoch;
		unsigned nb = (unsigned)((L / 16 + 255) / 256);
		nb = nb < 1 ? 1 : (nb > 64 ? 64 : nb);
		LFK_LAUNCH(c, k_halo_p2p, nb, 256, 0, c->arena, c->arena_peer[0], c->arena_peer[1], f + (size_t)nzl * L, f + L,
			f + (size_t)(nzl + 1) * L, f, L, c->arena_slot);
		return 0;
	}
	LFK_NCCL(c, g_nccl.GroupStart());
	if (c->rank + 1 < c->nranks) {
		LFK_NCCL(c, g_nccl.Send(f + (size_t)nzl * L, layer_elems, dt, c->rank + 1, comm, c->stream));
		LFK_NCCL(c, g_nccl.Recv(f + (size_t)(nzl + 1) * L, layer_elems, dt, c->rank + 1, comm, c->stream));
	}
	if (c->rank > 0) {
		LFK_NCCL(c, g_nccl.Send(f + L, layer_elems, dt, c->rank - 1, comm, c->stream));
		LFK_NCCL(c, g_nccl.Recv(f, layer_elems, dt, c->rank - 1, comm, c->stream));
	}
	LFK_NCCL(c, g_nccl.GroupEnd());
	return 0;
}

// my second-to-top owned layer (z0 + nzl - 2) goes to the upper rank, which sees it as layer z0 - 2
int lfkx_layer_below(lfk_ctx *c, const double *field, double *dst) {
	if (c->nranks == 1) { return 0; }
	PhaseTimer T(c, LFK_PHASE_EXCHANGE);
	ncclComm_t comm = (ncclComm_t)c->comm;
	const size_t L = (size_t)c->g.sxy;
	LFK_NCCL(c, g_nccl.GroupStart());
	if (c->rank + 1 < c->nranks) {
		LFK_NCCL(c, g_nccl.Send(field + (size_t)(c->g.nzl - 1) * L, L, ncclDouble, c->rank + 1, comm, c->stream));
	}
	if (c->rank > 0) {
		LFK_NCCL(c, g_nccl.Recv(dst, L, ncclDouble, c->rank - 1, comm, c->stream));
	}
	LFK_NCCL(c, g_nccl.GroupEnd());
	return 0;
}

int lfkx_halo_f64(lfk_ctx *c, double *field) {
	return halo_bytes(c, field, (size_t)c->g.sxy, c->g.nzl, ncclDouble, 8);
}
int lfkx_halo_f32(lfk_ctx *c, float *field, int nx, int ny, int nzl) {
	return halo_bytes(c, field, (size_t)nx * ny, nzl, ncclFloat, 4);
}
int lfkx_halo_u8(lfk_ctx *c, uint8_t *field) {
	return halo_bytes(c, field, (size_t)c->g.sxy, c->g.nzl, ncclUint8, 1);
}
int lfkx_allreduce_sum(lfk_ctx *c, double *d_vals, int n) {
	if (c->nranks == 1) { return 0; }
	LFK_NCCL(c, g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
	return 0;
}
int lfkx_allreduce_sum_f32(lfk_ctx *c, float *d_vals, int n) {
	if (c->nranks == 1) { return 0; }
	LFK_NCCL(c, g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)c->comm, c->stream));
	return 0;
}
int lfkx_allreduce_max(lfk_ctx *c, double *d_vals, int n) {
	if (c->nranks == 1) { return 0; }
	LFK_NCCL(c, g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
	return 0;
}

// =========================================================================================================
// Particle exchange (once per step, between advection and the cell sort).
//
// Rank r owns the cells z in [z0, z0 + nzl).  After advection every own particle is classified by its z cell zc:
//   zc <  z0 - 1        left the slab for good        -> sent DOWN, marked dead (dropped by the sort)
//   zc == z0 - 1        now belongs to the lower rank -> sent DOWN, kept here as a ghost copy (the ghost layer)
//   zc == z0            own, bottom boundary layer    -> a ghost copy is sent DOWN
//   zc == z0 + nzl - 1  own, top boundary layer       -> a ghost copy is sent UP
//   zc == z0 + nzl      now belongs to the upper rank -> sent UP, kept here as a ghost copy
//   zc >  z0 + nzl      left the slab for good        -> sent UP, marked dead
// What arrives is appended behind the own particles; the sort's keys then sort immigrants into owned cells and
// ghost copies into the ghost layers with no further distinction.  P2G and the position correction read the ghost
// layers' particles exactly like the reference reads the neighbouring cells (src/simulation.cpp:300-330, 572-600).
// The order of every message is the order of the sender's array (block counts -> scan -> offsets), so the result
// does not depend on timing.
// =========================================================================================================
#define XCH_THREADS 1024
#define XCH_FIELDS 15 // position, velocity, cx, cy, cz

__device__ __forceinline__ void xch_classify(const GridDesc &G, double z, int has_up, int has_dn, bool &up, bool &dn,
	bool &dead) {
	const int zc = cell_coord_clamped(z, G.off[2], G, G.nz);
	const int top = G.z0 + G.nzl;
	up = has_up && zc >= top - 1;
	dn = has_dn && zc <= G.z0;
	dead = zc > top || zc < G.z0 - 1;
}

__global__ void __launch_bounds__(XCH_THREADS) k_xch_count(GridDesc G, const double *__restrict__ pz,
	unsigned long long n, int has_up, int has_dn, uint32_t *__restrict__ cnt_up, uint32_t *__restrict__ cnt_dn) {
	unsigned long long i = (unsigned long long)blockIdx.x * XCH_THREADS + threadIdx.x;
	bool up = false, dn = false, dead = false;
	if (i < n) { xch_classify(G, pz[i], has_up, has_dn, up, dn, dead); }
	int nu = __syncthreads_count(up), nd = __syncthreads_count(dn);
	if (threadIdx.x == 0) {
		cnt_up[blockIdx.x] = (uint32_t)nu;
		cnt_dn[blockIdx.x] = (uint32_t)nd;
	}
}

// exclusive position of this thread among the threads of the block whose flag is set (deterministic)
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t *warp_tot /* [32] shared */) {
	const unsigned b = __ballot_sync(0xffffffffu, flag);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) { warp_tot[w] = (uint32_t)__popc(b); }
	__syncthreads();
	uint32_t before = 0;
	for (int k = 0; k < w; ++k) { before += warp_tot[k]; }
	return before + (uint32_t)__popc(b & ((1u << lane) - 1u));
}

// messages are field-major ([field][particle]): both the pack stores and the unpack loads are coalesced (the r1 layout,
// 15 doubles per particle record, wrote and read 8 bytes at a stride of 120)
__global__ void __launch_bounds__(XCH_THREADS) k_xch_pack(GridDesc G, ParticleSoA P, unsigned long long n, int has_up,
	int has_dn, const uint32_t *__restrict__ off_up, const uint32_t *__restrict__ off_dn, double *__restrict__ send_up,
	double *__restrict__ send_dn, size_t n_up, size_t n_dn) {
	__shared__ uint32_t wt[32];
	unsigned long long i = (unsigned long long)blockIdx.x * XCH_THREADS + threadIdx.x;
	bool up = false, dn = false, dead = false;
	if (i < n) { xch_classify(G, P.f[PF_PZ][i], has_up, has_dn, up, dn, dead); }
	const uint32_t ru = block_rank(up, wt), rd = block_rank(dn, wt);
	if (up) {
		double *rec = send_up + (size_t)(off_up[blockIdx.x] + ru);
		for (int f = 0; f < XCH_FIELDS; ++f) { rec[(size_t)f * n_up] = P.f[f][i]; }
	}
	if (dn) {
		double *rec = send_dn + (size_t)(off_dn[blockIdx.x] + rd);
		for (int f = 0; f < XCH_FIELDS; ++f) { rec[(size_t)f * n_dn] = P.f[f][i]; }
	}
	if (dead) { P.f[PF_PZ][i] = __longlong_as_double(0x7ff8000000000000ll); } // NaN: the sort drops it
}

// recv: [message of the lower neighbour: XCH_FIELDS x n_dn][message of the upper neighbour: XCH_FIELDS x n_up]
__global__ void k_xch_unpack(ParticleSoA P, unsigned long long at, const double *__restrict__ recv,
	unsigned long long n_dn, unsigned long long n_up, int with_old) {
	unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_dn + n_up) { return; }
	const bool lower = j < n_dn;
	const double *msg = lower ? recv : recv + n_dn * XCH_FIELDS;
	const unsigned long long cnt = lower ? n_dn : n_up, k = lower ? j : j - n_dn;
	for (int f = 0; f < XCH_FIELDS; ++f) { P.f[f][at + j] = msg[(size_t)f * cnt + k]; }
	if (with_old) { // old_position == position for a particle in flight between two steps
		for (int d = 0; d < 3; ++d) { P.f[PF_OX + d][at + j] = msg[(size_t)d * cnt + k]; }
	}
}

static int grow(lfk_ctx *c, double **buf, size_t *cap, size_t need) {
	if (need <= *cap) { return 0; }
	if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
	size_t ncap = need + need / 4 + 1024;
	LFK_CUDA(c, cudaMalloc((void**)buf, ncap * XCH_FIELDS * sizeof(double)));
	*cap = ncap;
	return 0;
}

int lfkx_exchange_particles(lfk_ctx *c, uint64_t *n_in) {
	*n_in = c->np;
	if (c->nranks == 1) { return 0; }
	PhaseTimer T(c, LFK_PHASE_EXCHANGE);
	const GridDesc &G = c->g;
	ncclComm_t comm = (ncclComm_t)c->comm;
	const int has_up = c->rank + 1 < c->nranks ? 1 : 0, has_dn = c->rank > 0 ? 1 : 0;
	const uint64_t n = c->np;
	const unsigned nb = lfk_blocks((long long)n, XCH_THREADS);
	// per-block counts, their scans (nb + 1 entries each): [cnt_up | cnt_dn | off_up | off_dn]
	const size_t need = 4 * ((size_t)nb + 1);
	if (need > c->xcnt_n) {
		if (c->xcnt) { cudaFree(c->xcnt); c->xcnt = nullptr; }
		LFK_CUDA(c, cudaMalloc((void**)&c->xcnt, need * sizeof(uint32_t)));
		c->xcnt_n = need;
	}
	uint32_t *cnt_up = c->xcnt, *cnt_dn = cnt_up + nb + 1, *off_up = cnt_dn + nb + 1, *off_dn = off_up + nb + 1;
	ParticleSoA V = lfk_own_view(c);
	uint32_t *hc = c->h_xcounts;
	hc[0] = hc[1] = hc[2] = hc[3] = 0;
	if (n > 0) {
		LFK_LAUNCH(c, k_xch_count, nb, XCH_THREADS, 0, G, V.f[PF_PZ], (unsigned long long)n, has_up, has_dn, cnt_up, cnt_dn);
		LFK_TRY(lfkp_exclusive_scan_u32(c, cnt_up, off_up, nb, 0));
		LFK_TRY(lfkp_exclusive_scan_u32(c, cnt_dn, off_dn, nb, 0));
		LFK_TRY(lfk_readback(c, hc + 0, off_up + nb, sizeof(uint32_t)));
		LFK_TRY(lfk_readback(c, hc + 1, off_dn + nb, sizeof(uint32_t)));
		LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	const uint32_t send_up = hc[0], send_dn = hc[1];
	LFK_TRY(grow(c, &c->xsend[0], &c->xsend_cap[0], send_up));
	LFK_TRY(grow(c, &c->xsend[1], &c->xsend_cap[1], send_dn));
	if (n > 0) {
		LFK_LAUNCH(c, k_xch_pack, nb, XCH_THREADS, 0, G, V, (unsigned long long)n, has_up, has_dn, off_up, off_dn,
			c->xsend[0], c->xsend[1], (size_t)send_up, (size_t)send_dn);
	}
	// message sizes
	LFK_CUDA(c, cudaMemcpyAsync(c->xcounts, hc, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
	LFK_CUDA(c, cudaMemsetAsync(c->xcounts + 2, 0, 2 * sizeof(uint32_t), c->stream));
	LFK_NCCL(c, g_nccl.GroupStart());
	if (has_up) {
		LFK_NCCL(c, g_nccl.Send(c->xcounts + 0, 1, ncclUint32, c->rank + 1, comm, c->stream));
		LFK_NCCL(c, g_nccl.Recv(c->xcounts + 2, 1, ncclUint32, c->rank + 1, comm, c->stream));
	}
	if (has_dn) {
		LFK_NCCL(c, g_nccl.Send(c->xcounts + 1, 1, ncclUint32, c->rank - 1, comm, c->stream));
		LFK_NCCL(c, g_nccl.Recv(c->xcounts + 3, 1, ncclUint32, c->rank - 1, comm, c->stream));
	}
	LFK_NCCL(c, g_nccl.GroupEnd());
	LFK_TRY(lfk_readback(c, hc + 2, c->xcounts + 2, 2 * sizeof(uint32_t)));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	const uint32_t recv_up = hc[2], recv_dn = hc[3];
	// payload: what the lower neighbour sent goes first (deterministic append order)
	const size_t nrecv = (size_t)recv_up + recv_dn;
	LFK_TRY(grow(c, &c->xrecv, &c->xrecv_cap, nrecv));
	LFK_NCCL(c, g_nccl.GroupStart());
	if (has_up) {
		if (send_up) { LFK_NCCL(c, g_nccl.Send(c->xsend[0], (size_t)send_up * XCH_FIELDS, ncclDouble, c->rank + 1, comm, c->stream)); }
		if (recv_up) {
			LFK_NCCL(c, g_nccl.Recv(c->xrecv + (size_t)recv_dn * XCH_FIELDS, (size_t)recv_up * XCH_FIELDS, ncclDouble,
				c->rank + 1, comm, c->stream));
		}
	}
	if (has_dn) {
		if (send_dn) { LFK_NCCL(c, g_nccl.Send(c->xsend[1], (size_t)send_dn * XCH_FIELDS, ncclDouble, c->rank - 1, comm, c->stream)); }
		if (recv_dn) { LFK_NCCL(c, g_nccl.Recv(c->xrecv, (size_t)recv_dn * XCH_FIELDS, ncclDouble, c->rank - 1, comm, c->stream)); }
	}
	LFK_NCCL(c, g_nccl.GroupEnd());
	if (nrecv > 0) {
		LFK_TRY(lfkp_reserve_particles(c, n + nrecv)); // may move the own particles to the front (first = 0)
		LFK_LAUNCH(c, k_xch_unpack, lfk_blocks((long long)nrecv, 256), 256, 0, c->P, (unsigned long long)(c->first + n),
			c->xrecv, (unsigned long long)recv_dn, (unsigned long long)recv_up, c->old_valid ? 1 : 0);
	}
	c->stats.exchanged_particles = (uint64_t)send_up + send_dn;
	*n_in = n + nrecv;
	return 0;
}
