// Multi-GPU plumbing: one context per rank owns a z-slab; neighbours exchange one ghost layer of cell data and
// the PCG scalars are all-reduced.  NCCL is resolved at run time with dlopen (libnccl.so.2 -- inside a PyTorch
// process that is torch's own copy), so single-GPU use has no NCCL dependency at all.
#include "lfk_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

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

int lfkx_init(lfk_ctx *c, const void *nccl_id128) {
	if (c->nranks == 1) { return 0; }
	LFK_REQUIRE(c, nccl_id128 != nullptr, LFK_E_INVALID, "nranks > 1 needs an NCCL unique id");
	LFK_REQUIRE(c, load_nccl(), LFK_E_NCCL, "libnccl.so.2 could not be loaded");
	ncclUniqueId id;
	memcpy(&id, nccl_id128, 128);
	ncclComm_t comm;
	LFK_NCCL(c, g_nccl.CommInitRank(&comm, c->nranks, id, c->rank));
	c->comm = comm;
	return 0;
}

int lfkx_destroy(lfk_ctx *c) {
	if (c->comm) {
		g_nccl.CommDestroy((ncclComm_t)c->comm);
		c->comm = nullptr;
	}
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
int lfkx_allreduce_max(lfk_ctx *c, double *d_vals, int n) {
	if (c->nranks == 1) { return 0; }
	LFK_NCCL(c, g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
	return 0;
}
