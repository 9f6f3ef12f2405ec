// Per-particle kernels: AoS<->SoA, cell keys + stable cell sort, advection, collisions, position correction,
// CFL reduction, grid-to-particle transfer.
//
// This file is compiled with --fmad=false: the reference's x86-64 build performs no FMA contraction, and the
// particle-motion code is branchy (DDA marching, clamps, truncations), so keeping plain IEEE mul/add makes cell
// indices bit-exact and positions reproducible against the CPU path.  All of these kernels are bandwidth bound;
// the split mul/add costs nothing measurable.
#include "lfk_internal.cuh"

#include <cstring>

// =========================================================================================================
// AoS (152-byte reference records) <-> SoA, staged through shared memory so that both sides are coalesced
// =========================================================================================================
#define AOS_WORDS 19 // 18 doubles + size_t
#define AOS_TILE 128

__global__ void __launch_bounds__(AOS_TILE) k_aos_to_soa(const unsigned long long *__restrict__ aos, ParticleSoA P,
	uint32_t *__restrict__ key, unsigned long long n, long long key_shift) {
	__shared__ unsigned long long sm[AOS_TILE * AOS_WORDS];
	unsigned long long base = (unsigned long long)blockIdx.x * AOS_TILE;
	unsigned long long cnt = n - base < AOS_TILE ? n - base : AOS_TILE;
	for (unsigned k = threadIdx.x; k < cnt * AOS_WORDS; k += AOS_TILE) {
		sm[k] = aos[base * AOS_WORDS + k];
	}
	__syncthreads();
	if (threadIdx.x < cnt) {
		const unsigned long long *rec = sm + threadIdx.x * AOS_WORDS; // stride 19 words: conflict-free (odd)
#pragma unroll
		for (int f = 0; f < PF_COUNT; ++f) {
			P.f[f][base + threadIdx.x] = __longlong_as_double((long long)rec[f]);
		}
		key[base + threadIdx.x] = (uint32_t)((long long)rec[18] + key_shift);
	}
}

__global__ void __launch_bounds__(AOS_TILE) k_soa_to_aos(unsigned long long *__restrict__ aos, ParticleSoA P,
	const uint32_t *__restrict__ key, unsigned long long n, int old_valid, long long key_shift) {
	__shared__ unsigned long long sm[AOS_TILE * AOS_WORDS];
	unsigned long long base = (unsigned long long)blockIdx.x * AOS_TILE;
	unsigned long long cnt = n - base < AOS_TILE ? n - base : AOS_TILE;
	if (threadIdx.x < cnt) {
		unsigned long long *rec = sm + threadIdx.x * AOS_WORDS;
		unsigned long long i = base + threadIdx.x;
#pragma unroll
		for (int f = 0; f < 15; ++f) {
			rec[f] = (unsigned long long)__double_as_longlong(P.f[f][i]);
		}
#pragma unroll
		for (int f = 0; f < 3; ++f) { // old_position == position unless materialised
			rec[15 + f] = (unsigned long long)__double_as_longlong(old_valid ? P.f[PF_OX + f][i] : P.f[PF_PX + f][i]);
		}
		rec[18] = (unsigned long long)((long long)key[i] - key_shift);
	}
	__syncthreads();
	for (unsigned k = threadIdx.x; k < cnt * AOS_WORDS; k += AOS_TILE) {
		aos[base * AOS_WORDS + k] = sm[k];
	}
}

__global__ void k_positions_interleave(double *__restrict__ xyz, const double *__restrict__ px,
	const double *__restrict__ py, const double *__restrict__ pz, unsigned long long n) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		xyz[3 * i] = px[i];
		xyz[3 * i + 1] = py[i];
		xyz[3 * i + 2] = pz[i];
	}
}

static ParticleSoA view_at(const lfk_ctx *c, uint64_t at) {
	ParticleSoA v = c->P;
	for (int f = 0; f < PF_COUNT; ++f) { v.f[f] += c->first + at; }
	return v;
}
int lfkp_aos_to_soa(lfk_ctx *c, const void *d_aos, uint64_t n, uint64_t at) {
	c->speed2_valid = false;
	if (n == 0) { return 0; }
	// raw_cell_index is a whole-grid raw index; device keys are local (one ghost layer below the slab)
	long long shift = ((long long)1 - c->g.z0) * c->g.sxy;
	LFK_LAUNCH(c, k_aos_to_soa, lfk_blocks((long long)n, AOS_TILE), AOS_TILE, 0,
		(const unsigned long long*)d_aos, view_at(c, at), c->key + c->first + at, (unsigned long long)n, shift);
	return 0;
}
int lfkp_soa_to_aos(lfk_ctx *c, void *d_aos, uint64_t n, uint64_t at) {
	if (n == 0) { return 0; }
	long long shift = ((long long)1 - c->g.z0) * c->g.sxy;
	LFK_LAUNCH(c, k_soa_to_aos, lfk_blocks((long long)n, AOS_TILE), AOS_TILE, 0,
		(unsigned long long*)d_aos, view_at(c, at), c->key + c->first + at, (unsigned long long)n, c->old_valid ? 1 : 0, shift);
	return 0;
}
int lfkp_positions_to_aos(lfk_ctx *c, double *d_xyz, uint64_t n) {
	if (n == 0) { return 0; }
	LFK_LAUNCH(c, k_positions_interleave, lfk_blocks((long long)n, 256), 256, 0,
		d_xyz, c->P.f[PF_PX] + c->first, c->P.f[PF_PY] + c->first, c->P.f[PF_PZ] + c->first, (unsigned long long)n);
	return 0;
}

// =========================================================================================================
// K1: cell keys (reference src/simulation.cpp:251-261) -- must be bit-exact: IEEE sub, IEEE div, truncation
// =========================================================================================================
__global__ void k_keys_hist(GridDesc G, const double *__restrict__ px, const double *__restrict__ py,
	const double *__restrict__ pz, uint32_t *__restrict__ key, uint32_t *__restrict__ slot,
	uint32_t *__restrict__ cnt, unsigned long long n) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	bool active = i < n;
	unsigned amask = __ballot_sync(0xffffffffu, active);
	if (!active) { return; }
	int x = cell_coord_clamped(px[i], G.off[0], G, G.nx);
	int y = cell_coord_clamped(py[i], G.off[1], G, G.ny);
	int z = cell_coord_clamped(pz[i], G.off[2], G, G.nz);
	int lz = z - G.z0 + 1;
	lz = lz < 0 ? 0 : (lz > G.nlz - 1 ? G.nlz - 1 : lz);
	uint32_t k = (uint32_t)(x + (long long)G.nx * (y + (long long)G.ny * lz));
	// multi-GPU: a particle that left for a neighbour beyond the ghost layer was marked (z = NaN) by the exchange; it
	// goes to the graveyard bin behind the last cell and drops off the end of the sorted array
	const double zz = pz[i];
	if (zz != zz) { k = (uint32_t)G.ncl; }
	key[i] = k;
	// warp-aggregated histogram: particles arrive nearly sorted, so a warp touches only a handful of cells
	unsigned peers = __match_any_sync(amask, k);
	int lane = threadIdx.x & 31;
	int leader = __ffs(peers) - 1;
	uint32_t basev = 0;
	if (lane == leader) {
		basev = atomicAdd(cnt + k, (uint32_t)__popc(peers));
	}
	basev = __shfl_sync(peers, basev, leader);
	slot[i] = basev + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// ---- exclusive scan over u32 (cell counts -> begin offsets) ------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t scan_load(const uint32_t *in, long long i, long long n, int from_flags) {
	if (i >= n) { return 0u; }
	uint32_t v = in[i];
	return from_flags ? (v > 0u ? 1u : 0u) : v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ in,
	uint32_t *__restrict__ tile_sum, long long n, int from_flags) {
	__shared__ uint32_t sm[SCAN_THREADS / 32];
	long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
	uint32_t s = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		s += scan_load(in, base + k, n, from_flags);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = s; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int k = 0; k < SCAN_THREADS / 32; ++k) { t += sm[k]; }
		tile_sum[blockIdx.x] = t;
	}
}

// single block: exclusive scan of the tile sums in place, total appended at [ntiles]
__global__ void __launch_bounds__(1024) k_scan_tile_offsets(uint32_t *__restrict__ tile_sum, long long ntiles) {
	__shared__ uint32_t warp_tot[32];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) { carry = 0; }
	__syncthreads();
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (long long base = 0; base < ntiles; base += 1024) {
		long long i = base + threadIdx.x;
		uint32_t v = i < ntiles ? tile_sum[i] : 0u, incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) { incl += t; }
		}
		if (lane == 31) { warp_tot[w] = incl; }
		__syncthreads();
		if (w == 0) {
			uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
				if (lane >= o) { ti += u; }
			}
			warp_tot[lane] = ti - t; // exclusive
		}
		__syncthreads();
		uint32_t excl = carry + warp_tot[w] + incl - v;
		if (i < ntiles) { tile_sum[i] = excl; }
		__syncthreads();
		if (threadIdx.x == 1023) { carry = excl + v; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { tile_sum[ntiles] = carry; }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ in,
	const uint32_t *__restrict__ tile_off, uint32_t *__restrict__ out, long long n, long long ntiles, int from_flags) {
	__shared__ uint32_t warp_tot[SCAN_THREADS / 32];
	long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		v[k] = scan_load(in, base + k, n, from_flags);
		s += v[k];
	}
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t incl = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) { incl += t; }
	}
	if (lane == 31) { warp_tot[w] = incl; }
	__syncthreads();
	uint32_t wbase = 0;
	for (int k = 0; k < w; ++k) { wbase += warp_tot[k]; }
	uint32_t run = tile_off[blockIdx.x] + wbase + incl - s;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		if (base + k < n) { out[base + k] = run; }
		run += v[k];
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { out[n] = tile_off[ntiles]; }
}

int lfkp_exclusive_scan_u32(lfk_ctx *c, const uint32_t *in, uint32_t *out, long long n, int from_flags) {
	long long ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (ntiles < 1) { ntiles = 1; }
	if ((size_t)(ntiles + 1) > c->scan_tmp_n) {
		if (c->scan_tmp) { cudaFree(c->scan_tmp); c->scan_tmp = nullptr; }
		LFK_CUDA(c, cudaMalloc(&c->scan_tmp, (size_t)(ntiles + 1) * sizeof(uint32_t)));
		c->scan_tmp_n = (size_t)(ntiles + 1);
	}
	LFK_LAUNCH(c, k_scan_tile_sums, (unsigned)ntiles, SCAN_THREADS, 0, in, c->scan_tmp, n, from_flags);
	LFK_LAUNCH(c, k_scan_tile_offsets, 1, 1024, 0, c->scan_tmp, ntiles);
	LFK_LAUNCH(c, k_scan_apply, (unsigned)ntiles, SCAN_THREADS, 0, in, c->scan_tmp, out, n, ntiles, from_flags);
	return 0;
}

// ---- K2: counting sort by cell, made stable by canonicalising the order inside each cell -------------------
// perm holds ABSOLUTE source indices (entry i of the input view is element first + i of the arrays)
__global__ void k_scatter_perm(const uint32_t *__restrict__ key, const uint32_t *__restrict__ slot,
	const uint32_t *__restrict__ begin, uint32_t *__restrict__ perm, unsigned long long n, uint32_t first) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		perm[begin[key[i]] + slot[i]] = first + (uint32_t)i;
	}
}

#define SMALL_CELL 32
#define BIG_CELL_LIMIT 16384
// The atomics above hand out in-cell slots in arrival order.  Sorting each cell's slice of `perm` by source
// index turns the counting sort into a STABLE sort by key => run-to-run deterministic particle order.
__global__ void k_sort_within_cells(const uint32_t *__restrict__ begin, uint32_t *__restrict__ perm,
	long long ncl, uint32_t *__restrict__ bigcells, unsigned *__restrict__ bigcount, unsigned bigcap) {
	long long cidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (cidx >= ncl) { return; }
	uint32_t b = begin[cidx], e = begin[cidx + 1], n = e - b;
	if (n < 2) { return; }
	if (n > SMALL_CELL) {
		if (n <= BIG_CELL_LIMIT) {
			unsigned k = atomicAdd(bigcount, 1u);
			if (k < bigcap) { bigcells[k] = (uint32_t)cidx; }
		}
		return; // beyond BIG_CELL_LIMIT (pathological pile-up) the order stays arrival order
	}
	uint32_t a[SMALL_CELL];
	for (uint32_t k = 0; k < n; ++k) { a[k] = perm[b + k]; }
	for (uint32_t k = 1; k < n; ++k) { // insertion sort; slices are short and almost sorted
		uint32_t v = a[k];
		int j = (int)k - 1;
		while (j >= 0 && a[j] > v) {
			a[j + 1] = a[j];
			--j;
		}
		a[j + 1] = v;
	}
	for (uint32_t k = 0; k < n; ++k) { perm[b + k] = a[k]; }
}

// one block per crowded cell: odd-even transposition sort in place
__global__ void __launch_bounds__(256) k_sort_big_cells(const uint32_t *__restrict__ begin,
	uint32_t *__restrict__ perm, const uint32_t *__restrict__ bigcells, const unsigned *__restrict__ bigcount,
	unsigned bigcap) {
	unsigned nbig = *bigcount < bigcap ? *bigcount : bigcap;
	for (unsigned bc = blockIdx.x; bc < nbig; bc += gridDim.x) {
		uint32_t cidx = bigcells[bc];
		uint32_t b = begin[cidx], n = begin[cidx + 1] - b;
		uint32_t *a = perm + b;
		for (uint32_t phase = 0; phase < n; ++phase) {
			for (uint32_t k = 2 * threadIdx.x + (phase & 1u); k + 1 < n; k += 2 * blockDim.x) {
				uint32_t lo = a[k], hi = a[k + 1];
				if (lo > hi) {
					a[k] = hi;
					a[k + 1] = lo;
				}
			}
			__syncthreads();
		}
	}
}

// dst[i] = src[perm[i]] for the field groups selected by `groups` (bit g <-> fields 3g .. 3g + 2: position, velocity,
// cx, cy, cz, old_position) and, with_key, the cell key
__global__ void __launch_bounds__(256) k_gather_particles(ParticleSoA dst, ParticleSoA src,
	uint32_t *__restrict__ key_dst, const uint32_t *__restrict__ key_src, const uint32_t *__restrict__ perm,
	unsigned long long n, unsigned groups, int with_key) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	const uint32_t s = perm[i];
	if (with_key) { key_dst[i] = key_src[s]; }
#pragma unroll
	for (int g = 0; g < 6; ++g) {
		if (groups & (1u << g)) {
			const double a = src.f[3 * g][s], b = src.f[3 * g + 1][s], c = src.f[3 * g + 2][s];
			dst.f[3 * g][i] = a;
			dst.f[3 * g + 1][i] = b;
			dst.f[3 * g + 2][i] = c;
		}
	}
}

// gathers the selected field groups through c->perm into the alternate buffers and makes those current
static int permute_groups(lfk_ctx *c, unsigned groups, bool with_key, uint64_t n) {
	if (n == 0 || (groups == 0 && !with_key)) { return 0; }
	LFK_LAUNCH(c, k_gather_particles, lfk_blocks((long long)n, 256), 256, 0, c->Palt, c->P, c->key_alt, c->key,
		c->perm, (unsigned long long)n, groups, with_key ? 1 : 0);
	for (int g = 0; g < 6; ++g) {
		if (groups & (1u << g)) {
			for (int f = 3 * g; f < 3 * g + 3; ++f) {
				double *t = c->P.f[f]; c->P.f[f] = c->Palt.f[f]; c->Palt.f[f] = t;
			}
		}
	}
	if (with_key) {
		uint32_t *kt = c->key; c->key = c->key_alt; c->key_alt = kt;
	}
	return 0;
}

#define GROUP_POS 1u
#define GROUP_VEL 2u
#define GROUP_C (4u | 8u | 16u)
#define GROUP_OLD 32u

// brings a velocity / c payload that a lean sort left in the pre-sort order (see lfkp_hash) into particle order
int lfkp_materialise_vc(lfk_ctx *c) {
	unsigned groups = (c->v_deferred ? GROUP_VEL : 0u) | (c->c_deferred ? GROUP_C : 0u);
	c->v_deferred = false;
	c->c_deferred = false;
	return permute_groups(c, groups, false, c->ntot);
}
int lfkp_permute_c(lfk_ctx *c) {
	c->c_deferred = false;
	return permute_groups(c, GROUP_C, false, c->ntot);
}

// K1 + K2.  lean: only the positions (and keys) are physically permuted.  The velocity -- and for APIC the c rows --
// stay where they are and are read through `perm` by the one kernel that still needs them (P2G; FLIP's G2P), because
// G2P overwrites them in the new order anyway: that removes 192 of the 240 B/particle the full permutation moves.
int lfkp_hash(lfk_ctx *c, bool lean) {
	PhaseTimer T(c, LFK_PHASE_SORT);
	const GridDesc &G = c->g;
	const bool multi = c->nranks > 1;
	LFK_TRY(lfkp_materialise_vc(c)); // a pending permutation cannot be composed with a new one
	// entries to sort: the own particles, plus -- multi-GPU -- what the neighbours sent (immigrants and ghost copies of
	// their boundary layers), appended behind them
	uint64_t n = c->np;
	if (multi) { LFK_TRY(lfkx_exchange_particles(c, &n)); }
	const size_t nbins = (size_t)G.ncl + 1; // + the graveyard bin
	LFK_CUDA(c, cudaMemsetAsync(c->cnt, 0, nbins * sizeof(uint32_t), c->stream));
	if (n > 0) {
		LFK_LAUNCH(c, k_keys_hist, lfk_blocks((long long)n, 256), 256, 0, G, c->P.f[PF_PX] + c->first,
			c->P.f[PF_PY] + c->first, c->P.f[PF_PZ] + c->first, c->key + c->first, c->slot, c->cnt,
			(unsigned long long)n);
	}
	LFK_TRY(lfkp_exclusive_scan_u32(c, c->cnt, c->begin, (long long)nbins, 0));
	if (n > 0) {
		LFK_LAUNCH(c, k_scatter_perm, lfk_blocks((long long)n, 256), 256, 0, c->key + c->first, c->slot, c->begin,
			c->perm, (unsigned long long)n, (uint32_t)c->first);
		LFK_CUDA(c, cudaMemsetAsync(c->bigcount, 0, sizeof(unsigned), c->stream));
		LFK_LAUNCH(c, k_sort_within_cells, lfk_blocks(G.ncl, 128), 128, 0, c->begin, c->perm, G.ncl, c->bigcells,
			c->bigcount, c->bigcap);
		// crowded cells (> 32 particles) are rare; a small grid strides over however many there are
		if (n > SMALL_CELL) {
			LFK_LAUNCH(c, k_sort_big_cells, 296, 256, 0, c->begin, c->perm, c->bigcells, c->bigcount, c->bigcap);
		}
		unsigned groups = GROUP_POS | (c->old_valid ? GROUP_OLD : 0u);
		if (lean) {
			c->v_deferred = true;
			c->c_deferred = c->prm.method == LFK_METHOD_APIC;
			if (!c->c_deferred) { groups |= GROUP_C; }
		} else {
			groups |= GROUP_VEL | GROUP_C;
		}
		LFK_TRY(permute_groups(c, groups, true, n));
	}
	if (multi) { // own range of the sorted array = the owned layers; ghosts sit before / behind it, the dead at the end
		const uint32_t *src[3] = { c->begin + G.sxy, c->begin + G.sxy * (G.nzl + 1), c->begin + G.ncl };
		for (int k = 0; k < 3; ++k) {
			LFK_TRY(lfk_readback(c, c->h_xcounts + 4 + k, src[k], sizeof(uint32_t)));
		}
		LFK_CUDA(c, cudaStreamSynchronize(c->stream));
		c->first = c->h_xcounts[4];
		c->np = (uint64_t)c->h_xcounts[5] - c->h_xcounts[4];
		c->ntot = c->h_xcounts[6];
		// from here on `cnt` means "own particles per cell" (solver unknowns, fluid-cell list); the ghost layers'
		// particles stay reachable through `begin`
		LFK_CUDA(c, cudaMemsetAsync(c->cnt, 0, (size_t)G.sxy * sizeof(uint32_t), c->stream));
		LFK_CUDA(c, cudaMemsetAsync(c->cnt + (G.ncl - G.sxy), 0, (size_t)G.sxy * sizeof(uint32_t), c->stream));
	} else {
		c->first = 0;
		c->ntot = n;
	}
	c->table_valid = true;
	c->keys_valid = true;
	c->ordinal_valid = false;
	c->system_valid = false;
	return 0;
}

// =========================================================================================================
// A1: advection (reference src/simulation.cpp:240-248)
// =========================================================================================================
struct MotionParams {
	double lo[3], hi[3]; // advect clamp corners
	double gmin[3], gmax[3]; // correct clamp corners
	double skin, skin_max;
	double dt, corr_factor, re2, inv_re2;
};

static MotionParams motion_params(const lfk_ctx *c, double dt) {
	MotionParams m;
	const GridDesc &G = c->g;
	const double size[3] = { (double)G.nx, (double)G.ny, (double)G.nz };
	double skin = c->prm.boundary_skin_width;
	for (int d = 0; d < 3; ++d) {
		m.lo[d] = G.off[d] + skin;                      // grid_offset + skin_width
		m.hi[d] = G.h * size[d] + G.off[d] - skin;       // cell_size * size + grid_offset - skin_width
		m.gmin[d] = G.off[d];
		m.gmax[d] = G.off[d] + size[d] * G.h;            // grid_offset + size * cell_size
	}
	m.skin = skin;
	m.skin_max = G.h - skin;
	m.dt = dt;
	double re = G.h / sqrt(2.0);
	m.corr_factor = dt * c->prm.correction_stiffness * re;
	m.re2 = re * re;
	m.inv_re2 = 1.0 / m.re2;
	return m;
}

__device__ __forceinline__ void advect_one(const MotionParams &M, double *p, const double *v) {
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		p[d] = dclamp_std(p[d] + v[d] * M.dt, M.lo[d], M.hi[d]);
	}
}

// =========================================================================================================
// A2: collisions = DDA march old -> new through the cell types + skin push-out
// (reference src/simulation.cpp:612-683, include/fluid/data_structures/grid.h:140-209)
// =========================================================================================================
__device__ __forceinline__ bool cell_is_free(const GridDesc &G, const uint8_t *__restrict__ typ, int x, int y, int z) {
	if (x < 0 || y < 0 || z < 0 || x >= G.nx || y >= G.ny || z >= G.nz) {
		return false;
	}
	int lz = z - G.z0 + 1;
	if (lz < 0 || lz >= G.nlz) { // beyond this rank's halo: cannot be decided here, treat as free (see exchange.cu)
		return true;
	}
	return typ[x + (long long)G.nx * (y + (long long)G.ny * lz)] != LFK_CELL_SOLID;
}

__device__ void collide_one(const GridDesc &G, const MotionParams &M, const uint8_t *__restrict__ typ,
	double *from, double *to) {
	const double h = G.h;
	for (int j = 0; j < 3; ++j) {
		bool into_wall = false;
		double gf[3], gt[3], inv[3], normal[3], t[3];
		int cur[3], tc[3], adv[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			gf[d] = div_h(from[d] - G.off[d], G);
			gt[d] = div_h(to[d] - G.off[d], G);
			cur[d] = (int)floor(gf[d]);
			tc[d] = (int)floor(gt[d]);
		}
		if (cur[0] == tc[0] && cur[1] == tc[1] && cur[2] == tc[2]) {
			break; // no cell boundary is crossed: the march below would not take a single step
		}
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			double diff = gt[d] - gf[d];
			int face;
			if (diff > 0.0) {
				adv[d] = 1;
				face = 1;
			} else {
				adv[d] = -1;
				face = 0;
			}
			inv[d] = 1.0 / fabs(diff);
			normal[d] = -(double)adv[d];
			t[d] = fabs((double)(cur[d] + face) - gf[d]) * inv[d];
		}
		while (cur[0] != tc[0] || cur[1] != tc[1] || cur[2] != tc[2]) {
			int mc = 0;
			double mint = 2.0;
#pragma unroll
			for (int d = 0; d < 3; ++d) {
				if (t[d] < mint) {
					mint = t[d];
					mc = d;
				}
			}
			if (!(mint <= 1.0)) {
				break;
			}
			// (dynamic indexing avoided: select by mc)
			if (mc == 0) { cur[0] += adv[0]; } else if (mc == 1) { cur[1] += adv[1]; } else { cur[2] += adv[2]; }
			if (!cell_is_free(G, typ, cur[0], cur[1], cur[2])) {
				double offv[3] = { to[0] - from[0], to[1] - from[1], to[2] - from[2] };
				double nrm[3] = { 0.0, 0.0, 0.0 };
				double tm = mc == 0 ? t[0] : (mc == 1 ? t[1] : t[2]);
				double nm = mc == 0 ? normal[0] : (mc == 1 ? normal[1] : normal[2]);
				if (mc == 0) { nrm[0] = nm; } else if (mc == 1) { nrm[1] = nm; } else { nrm[2] = nm; }
				double dotv = 0.0;
				dotv += offv[0] * nrm[0];
				dotv += offv[1] * nrm[1];
				dotv += offv[2] * nrm[2];
				double tt = tm + M.skin / dotv;
				tt = dmax_std(tt, 0.0);
#pragma unroll
				for (int d = 0; d < 3; ++d) {
					from[d] = tt * to[d] + (1.0 - tt) * from[d];
				}
				if (mc == 0) { to[0] = from[0]; } else if (mc == 1) { to[1] = from[1]; } else { to[2] = from[2]; }
				into_wall = true;
				break;
			}
			if (mc == 0) { t[0] += inv[0]; } else if (mc == 1) { t[1] += inv[1]; } else { t[2] += inv[2]; }
		}
		if (!into_wall) {
			break;
		}
	}
	// skin push-out: cell index and in-cell position are computed once, before the per-axis pushes
	double cp[3];
	int ci[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		double gp = to[d] - G.off[d];
		unsigned long long idx = (unsigned long long)div_h(gp, G);
		ci[d] = (int)(idx < 0x7fffffffull ? idx : 0x7fffffffull);
		cp[d] = gp - (double)idx * h;
	}
	const int size[3] = { G.nx, G.ny, G.nz };
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		int q[3] = { ci[0], ci[1], ci[2] };
		if (cp[d] < M.skin) {
			q[d] = ci[d] - 1;
			if (ci[d] == 0 || !cell_is_free(G, typ, q[0], q[1], q[2])) {
				to[d] += M.skin - cp[d];
			}
		}
		if (cp[d] > M.skin_max) {
			q[d] = ci[d] + 1;
			if (ci[d] + 1 >= size[d] || !cell_is_free(G, typ, q[0], q[1], q[2])) {
				to[d] += M.skin_max - cp[d];
			}
		}
	}
}

__global__ void k_advect(MotionParams M, ParticleSoA P, unsigned long long n) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	double p[3] = { P.f[PF_PX][i], P.f[PF_PY][i], P.f[PF_PZ][i] };
	double v[3] = { P.f[PF_VX][i], P.f[PF_VY][i], P.f[PF_VZ][i] };
	advect_one(M, p, v);
	P.f[PF_PX][i] = p[0];
	P.f[PF_PY][i] = p[1];
	P.f[PF_PZ][i] = p[2];
}

__global__ void k_collide(GridDesc G, MotionParams M, ParticleSoA P, const uint8_t *__restrict__ typ,
	unsigned long long n, int old_valid) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	double to[3] = { P.f[PF_PX][i], P.f[PF_PY][i], P.f[PF_PZ][i] };
	double from[3] = { to[0], to[1], to[2] };
	if (old_valid) {
		from[0] = P.f[PF_OX][i];
		from[1] = P.f[PF_OY][i];
		from[2] = P.f[PF_OZ][i];
	}
	collide_one(G, M, typ, from, to);
	P.f[PF_PX][i] = to[0];
	P.f[PF_PY][i] = to[1];
	P.f[PF_PZ][i] = to[2];
}

// advect + collide in one pass: old_position is the pre-advection position held in registers.  (Queueing the particles
// that cross a cell boundary in shared memory and marching them densely packed was measured: 3.26 ms against 3.02 ms
// at 256^3, r2r sweep -- the block barrier and the queue cost more than the divergence of the march.)
#ifndef ADV_THREADS
#define ADV_THREADS 128
#endif
__global__ void k_advect_collide(GridDesc G, MotionParams M, ParticleSoA P, const uint8_t *__restrict__ typ,
	unsigned long long n) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	double from[3] = { P.f[PF_PX][i], P.f[PF_PY][i], P.f[PF_PZ][i] };
	double v[3] = { P.f[PF_VX][i], P.f[PF_VY][i], P.f[PF_VZ][i] };
	double to[3] = { from[0], from[1], from[2] };
	advect_one(M, to, v);
	collide_one(G, M, typ, from, to);
	P.f[PF_PX][i] = to[0];
	P.f[PF_PY][i] = to[1];
	P.f[PF_PZ][i] = to[2];
}

static int materialise_old(lfk_ctx *c) {
	if (!c->old_valid && c->np > 0) {
		for (int d = 0; d < 3; ++d) {
			LFK_CUDA(c, cudaMemcpyAsync(c->P.f[PF_OX + d] + c->first, c->P.f[PF_PX + d] + c->first, c->np * sizeof(double),
				cudaMemcpyDeviceToDevice, c->stream));
		}
	}
	c->old_valid = true;
	return 0;
}

int lfkp_advect(lfk_ctx *c, double dt) {
	PhaseTimer T(c, LFK_PHASE_ADVECT_COLLIDE);
	LFK_TRY(lfkp_materialise_vc(c));
	LFK_TRY(materialise_old(c));
	if (c->np == 0) { return 0; }
	LFK_LAUNCH(c, k_advect, lfk_blocks((long long)c->np, 256), 256, 0, motion_params(c, dt), lfk_own_view(c),
		(unsigned long long)c->np);
	c->table_valid = false;
	return 0;
}

int lfkp_collide(lfk_ctx *c) {
	PhaseTimer T(c, LFK_PHASE_ADVECT_COLLIDE);
	if (c->np > 0) {
		LFK_LAUNCH(c, k_collide, lfk_blocks((long long)c->np, 128), 128, 0, c->g, motion_params(c, 0.0), lfk_own_view(c), c->typ,
			(unsigned long long)c->np, c->old_valid ? 1 : 0);
	}
	c->old_valid = false; // old_position = position (reference src/simulation.cpp:57-59,115-117)
	return 0;
}

int lfkp_advect_collide(lfk_ctx *c, double dt) {
	PhaseTimer T(c, LFK_PHASE_ADVECT_COLLIDE);
	if (c->old_valid) { // a host upload left old != position: honour it with the unfused pair
		LFK_TRY(lfkp_advect(c, dt));
		return lfkp_collide(c);
	}
	LFK_TRY(lfkp_materialise_vc(c));
	if (c->np > 0) {
		LFK_LAUNCH(c, k_advect_collide, lfk_blocks((long long)c->np, ADV_THREADS), ADV_THREADS, 0, c->g, motion_params(c, dt),
			lfk_own_view(c), c->typ, (unsigned long long)c->np);
	}
	c->table_valid = false;
	return 0;
}

// =========================================================================================================
// A3: position correction (reference src/simulation.cpp:562-610)
// =========================================================================================================
#ifndef CT_KICK_INLINE
#define CT_KICK_INLINE __forceinline__
#endif
__device__ CT_KICK_INLINE void degenerate_kick(const double *p, const double *o, double *out3) {
	unsigned long long s = 0x9e3779b97f4a7c15ull;
#pragma unroll
	for (int k = 0; k < 3; ++k) { s = mix64(s ^ (unsigned long long)__double_as_longlong(p[k])); }
#pragma unroll
	for (int k = 0; k < 3; ++k) { s = mix64(s ^ (unsigned long long)__double_as_longlong(o[k])); }
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		s = mix64(s + 0x9e3779b97f4a7c15ull);
		out3[d] = (double)(s >> 11) * (2.0 / 9007199254740992.0) - 1.0;
	}
}

// A block owns a tile of CT_TY x CT_TZ rows x CT_LX cells and stages every particle of the tile plus its one-cell halo
// ONCE in shared memory as {x, y, z, |q|^2} in fp32, in cell units about the tile centre.
// Phase 1 (fp32, conservative pre-filter): per neighbouring row only the cells the kernel radius re = h / sqrt(2) can
// reach are scanned (re^2 - dy_min^2 - dz_min^2 leaves an x window of 1 .. 3 cells); with n = -2 r the test
// |r - q|^2 < 1/2 + margin is  w + n.q < T,  T = 1/2 + margin - |r|^2: three FFMA and a compare; hits are bits of a per-row
// mask with compile-time positions, one {mask, index of the window's first particle} record per 32 candidates.
// Every staged row is followed by CT_PAD entries that can never hit, so the 4-wide groups need no tail handling.
// Rounding: |coordinates| <= 17.1 cells (CT_LX <= 32), every intermediate below 620 with an error below 4e-5; the margin of 2e-3 cells^2
// covers their sum 10x over, so the filter never drops a true neighbour.
// Phase 2 (fp64): the recorded candidates in staging order (= the reference's order: rows by z then y, cells by x,
// particles in sorted order), evaluated from the original positions => the result is the plain fp64 loop's, bit for bit.
// Measured alternatives (256^3, 27.6 ms for this kernel): own particles grouped by (y, z) reach class 38.5 ms, 8-wide
// guarded groups 28.3 ms, packed f32x2 tests 42.6 ms, scalar d^2 test 30.3 ms (r2a sweep); fp64 positions staged in
// shared memory as well, one 1024-thread block per SM: 35.2 ms (r2b) -- two resident blocks per SM that overlap each
// other's staging and tails are worth more than the L1 / L2 latency of phase 2; a warp-cooperative re-mapping (fp64-only
// staging, a warp scans one own cell's 27-cell neighbourhood with lane <-> candidate and ballots, then lane <-> own
// particle evaluates from shared memory): bit-identical but 66.3 ms -- 277 warp instructions per particle against 147
// here (per-slot guards, ballot bookkeeping and the candidate-index searches cost more than the divergence they remove;
// profiles/r2l_coop_ncu_summary.txt, git history of r2 holds the kernel); two or three neighbour positions in flight in
// phase 2 instead of one: 27.7 / 28.2 ms against 27.3 (r2q sweep).
// Tile geometry (r3c-r3f sweeps at 256^3, tools/build_variant.py + tools/gpu_variants.sh; ms for this kernel):
//   threads x blocks/SM, CT_LX, CT_CAP      L1 left     ms
//   512 x 2, 32, 5120 (rounds 1-2)           60 KB     26.98
//   512 x 2, 32, 4960                        92 KB     26.34   (2 x 82 KB of shared memory fit the 164 KB carve-out)
//   384 x 2, 32, 5120 (68 registers)         60 KB     29.59   (24 warps per SM instead of 32)
//   256 x 4, 16, 2688                        60 KB     26.27   <- production
//   256 x 4, 16, 2432                        92 KB     25.96   (3 % headroom over a uniform 8 per cell: too tight)
//   256 x 4, 16, 3072                        28 KB     27.16
//   320 x 3, 20, 3200                        92 KB     26.44
//   256 x 4,  8, 1536                       124 KB     26.97   (x halo 10 / 8)
//   512 x 2, 16, 2688                       156 KB     28.30   (one pass per thread, coarse blocks)
//   128 x 8,  8, 1536                        28 KB     29.21
// i.e. 32 resident warps are needed, smaller blocks overlap each other's staging and tails better, and phase 2's
// neighbour loads want the L1 that the shared-memory carve-out leaves (about 0.3 ms per 32 KB).
#ifndef CT_LX
#define CT_LX 16
#endif
#define CT_TY 2
#define CT_TZ 2
#define CT_SY (CT_TY + 2)
#define CT_SZ (CT_TZ + 2)
#define CT_ROWS (CT_SY * CT_SZ)
#define CT_OWN (CT_TY * CT_TZ)
#ifndef CT_CAP
#define CT_CAP 2688           // staged particles per tile (42 KB; a uniform 8 per cell stages 2352); denser tiles take the global-memory path
#endif
#ifndef CT_THREADS
#define CT_THREADS 256
#endif
#ifndef CT_BLOCKS
#define CT_BLOCKS 4 // resident blocks per SM the register budget is set for
#endif
#define CT_PAD 3
#define CT_LIST 16
#define CT_MARGIN 2e-3f

// 1 / sqrt(x) for x in the normal range: hardware seed (rel. error 2^-22) + two Newton steps (~1 ulp)
__device__ __forceinline__ double fast_rsqrt(double x) {
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	const double hx = 0.5 * x;
	y = fma(y, fma(-hx * y, y, 0.5), y);
	y = fma(y, fma(-hx * y, y, 0.5), y);
	return y;
}

__device__ __forceinline__ void pair_exact(const MotionParams &M, const double *p, const double *o, double &sx,
	double &sy, double &sz) {
	double ox = p[0] - o[0], oy = p[1] - o[1], oz = p[2] - o[2];
	// explicit fused multiply-adds (the file is compiled --fmad=false for the bit-exact stages): 11 fp64 instructions
	// less per pair than the separate multiplies and adds; the summation order is the reference's
	const double sq = fma(oz, oz, fma(oy, oy, ox * ox));
	if (sq < 1e-12) {
		double kick[3];
		degenerate_kick(p, o, kick);
		sx += kick[0];
		sy += kick[1];
		sz += kick[2];
	} else {
		// reference: kernel = (1 - r^2 / re^2)^3, spring += kernel / sqrt(r^2) * offset.  The two IEEE divisions and
		// the IEEE square root cost ~200 instructions per pair; the reciprocal multiply and the Newton rsqrt (~1 ulp)
		// cost ~25 and move the corrected position by < 1e-16 cells -- positions are tolerance-checked (1e-12).
		const double kl = fma(-sq, M.inv_re2, 1.0);
		if (kl > 0.0) {
			const double sc = kl * kl * kl * fast_rsqrt(sq);
			sx = fma(sc, ox, sx);
			sy = fma(sc, oy, sy);
			sz = fma(sc, oz, sz);
		}
	}
}

// the plain fp64 neighbourhood loop (over-full tiles, particles outside their key cell, crowded rows): the reference's
// loop literally (src/simulation.cpp:572-600)
__device__ void spring_global(const GridDesc &G, const MotionParams &M, const double *__restrict__ px,
	const double *__restrict__ py, const double *__restrict__ pz, const uint32_t *__restrict__ begin,
	unsigned long long i, const double *p, double &sx, double &sy, double &sz) {
	int lo[3], hi[3];
	const int size[3] = { G.nx, G.ny, G.nz };
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		unsigned long long ci = (unsigned long long)div_h(p[d] - G.off[d], G);
		long long cl = ci > 0x7fffffffull ? 0x7fffffffll : (long long)ci;
		lo[d] = (int)(cl < 1 ? 0 : cl - 1);
		long long h2 = cl + 2;
		hi[d] = (int)(h2 < size[d] ? h2 : size[d]);
	}
	for (int cz = lo[2]; cz < hi[2]; ++cz) {
		int lz = cz - G.z0 + 1;
		if (lz < 0 || lz >= G.nlz) { continue; }
		for (int cy = lo[1]; cy < hi[1]; ++cy) {
			if (lo[0] >= hi[0]) { continue; }
			long long row = (long long)G.nx * (cy + (long long)G.ny * lz);
			uint32_t qb = begin[row + lo[0]], qe = begin[row + hi[0]];
			for (uint32_t q = qb; q < qe; ++q) {
				if (q == i) { continue; }
				double o[3] = { px[q], py[q], pz[q] };
				pair_exact(M, p, o, sx, sy, sz);
			}
		}
	}
}

template <bool COLLIDE> __global__ void __launch_bounds__(CT_THREADS, CT_BLOCKS) k_correct_tile(GridDesc G, MotionParams M,
	ParticleSoA P, double *__restrict__ nx_, double *__restrict__ ny_, double *__restrict__ nz_,
	const uint32_t *__restrict__ begin, const uint8_t *__restrict__ typ) {
	extern __shared__ float4 stage[];                            // [CT_CAP] fp32 scan entries
	__shared__ uint32_t rowstart[CT_ROWS], rowoff[CT_ROWS + 1], cellbeg[CT_ROWS][CT_LX + 3];
	__shared__ uint32_t ownbeg[CT_OWN], ownpre[CT_OWN + 1];
	const double *__restrict__ px = P.f[PF_PX], *__restrict__ py = P.f[PF_PY], *__restrict__ pz = P.f[PF_PZ];
	const int x0 = blockIdx.x * CT_LX, y0 = blockIdx.y * CT_TY, lz0 = blockIdx.z * CT_TZ + 1;
	const int tid = threadIdx.x;

	// ---- table of the staged rows: row r = (lz0 - 1 + r / CT_SY, y0 - 1 + r % CT_SY), cells x0 - 1 .. x0 + CT_LX ----
	for (int e = tid; e < CT_ROWS * (CT_LX + 3); e += CT_THREADS) {
		int r = e / (CT_LX + 3), k = e % (CT_LX + 3);
		int y = y0 - 1 + r % CT_SY, lz = lz0 - 1 + r / CT_SY;
		uint32_t v = 0;
		if (y >= 0 && y < G.ny && lz >= 0 && lz < G.nlz) {
			long long row = (long long)G.nx * (y + (long long)G.ny * lz);
			int xa = x0 - 1 < 0 ? 0 : x0 - 1;
			int xk = x0 - 1 + k;
			xk = xk < 0 ? 0 : (xk > G.nx ? G.nx : xk);
			uint32_t base = begin[row + xa];
			v = begin[row + xk] - base;
			if (k == 0) { rowstart[r] = base; }
		} else if (k == 0) {
			rowstart[r] = 0;
		}
		cellbeg[r][k] = v;
	}
	__syncthreads();
	if (tid == 0) {
		uint32_t acc = 0;
		for (int r = 0; r < CT_ROWS; ++r) { // every row is followed by CT_PAD never-hit entries
			rowoff[r] = acc;
			acc += cellbeg[r][CT_LX + 2] + CT_PAD;
		}
		rowoff[CT_ROWS] = acc;
		uint32_t oacc = 0;
		for (int o = 0; o < CT_OWN; ++o) { // own rows: cells x0 .. x0 + LX - 1 (clipped) of the inner rows
			int r = (o / CT_TY + 1) * CT_SY + (o % CT_TY + 1);
			int y = y0 + o % CT_TY, lz = lz0 + o / CT_TY;
			uint32_t nown = 0;
			if (y < G.ny && lz <= G.nzl) {
				nown = cellbeg[r][CT_LX + 1] - cellbeg[r][1];
			}
			ownbeg[o] = rowstart[r] + cellbeg[r][1];
			ownpre[o] = oacc;
			oacc += nown;
		}
		ownpre[CT_OWN] = oacc;
	}
	__syncthreads();
	const uint32_t nown_total = ownpre[CT_OWN];
	if (nown_total == 0) { return; }
	const uint32_t staged = rowoff[CT_ROWS];
	const bool use_stage = staged <= CT_CAP;
	// tile centre, fp64; the scan entries are relative to it, in cells
	const double ctr[3] = { G.off[0] + ((double)x0 + 0.5 * CT_LX) * G.h, G.off[1] + ((double)y0 + 0.5 * CT_TY) * G.h,
		G.off[2] + ((double)(lz0 - 1 + G.z0) + 0.5 * CT_TZ) * G.h };
	if (use_stage) {
		for (uint32_t e = tid; e < staged; e += CT_THREADS) {
			int r = 0; // row of staged entry e (binary search over the 16 row offsets)
			if (e >= rowoff[8]) { r = 8; }
			if (e >= rowoff[r + 4]) { r += 4; }
			if (e >= rowoff[r + 2]) { r += 2; }
			if (e >= rowoff[r + 1]) { r += 1; }
			const uint32_t j = e - rowoff[r];
			float4 f = make_float4(0.f, 0.f, 0.f, 1e30f); // padding: never within reach
			if (j < cellbeg[r][CT_LX + 2]) {
				const uint32_t q = rowstart[r] + j;
				f.x = (float)((px[q] - ctr[0]) * G.inv_h);
				f.y = (float)((py[q] - ctr[1]) * G.inv_h);
				f.z = (float)((pz[q] - ctr[2]) * G.inv_h);
				f.w = __fmaf_rn(f.z, f.z, __fmaf_rn(f.y, f.y, f.x * f.x));
			}
			stage[e] = f;
		}
	}
	__syncthreads();

	for (uint32_t t = tid; t < nown_total; t += CT_THREADS) {
		int o = 0;
#pragma unroll
		for (int k = 1; k < CT_OWN; ++k) {
			if (t >= ownpre[k]) { o = k; }
		}
		const uint32_t within = t - ownpre[o];
		const unsigned long long i = (unsigned long long)ownbeg[o] + within;
		const int oy = o % CT_TY, oz = o / CT_TY;
		const int rown = (oz + 1) * CT_SY + (oy + 1);
		const uint32_t so = rowoff[rown] + cellbeg[rown][1] + within; // this particle's own staged entry
		double p[3] = { px[i], py[i], pz[i] };
		double sx = 0.0, sy = 0.0, sz = 0.0;
		// unclamped cell and in-cell fraction (compute_cell_index)
		long long ci[3];
		float fr[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			double f = div_h(p[d] - G.off[d], G);
			unsigned long long u = (unsigned long long)f;
			ci[d] = u > 0x7fffffffull ? 0x7fffffffll : (long long)u;
			fr[d] = (float)(f - (double)u);
		}
		const bool in_tile = use_stage && ci[0] >= x0 && ci[0] < x0 + CT_LX && ci[0] < G.nx && ci[1] == y0 + oy &&
			ci[2] == lz0 + oz - 1 + G.z0;
		if (!in_tile) {
			spring_global(G, M, px, py, pz, begin, i, p, sx, sy, sz);
		} else {
			const float4 self = stage[so];
			const float nrx = -2.f * self.x, nry = -2.f * self.y, nrz = -2.f * self.z;
			// |r - q|^2 < 1/2 + margin  <=>  w + n.q < T
			const float T = (0.5f + CT_MARGIN) - self.w;
			const int kown = (int)(ci[0] - x0) + 1; // index of the particle's own cell in cellbeg[r][]
			// Phase 1: fp32 scan of the staged candidates; per 32 candidates of a row window one {hit mask, staged index
			// of the first candidate} record
			uint2 rec[CT_LIST];
			int nr = 0;
			for (int dz = -1; dz <= 1; ++dz) {
				// distance (in cells) from the particle to the nearest point of the neighbouring layer / row
				const float zmin = dz < 0 ? fr[2] : (dz > 0 ? 1.f - fr[2] : 0.f);
				const float remz = 0.501f - zmin * zmin; // (re / h)^2 = 1/2, plus a margin
				if (remz <= 0.f) { continue; }
				for (int dy = -1; dy <= 1; ++dy) {
					const float ymin = dy < 0 ? fr[1] : (dy > 0 ? 1.f - fr[1] : 0.f);
					const float rem = remz - ymin * ymin;
					if (rem <= 0.f) { continue; }
					const float xr = sqrtf(rem); // reach along x within this row: < 0.708 cells
					const int klo = kown + (fr[0] - xr < 0.f ? -1 : 0);
					const int khi = kown + (fr[0] + xr >= 1.f ? 2 : 1);
					const int r = (oz + 1 + dz) * CT_SY + (oy + 1 + dy);
					const uint32_t w0 = cellbeg[r][klo];
					uint32_t s = rowoff[r] + w0;
					const uint32_t s1 = rowoff[r] + cellbeg[r][khi];
					uint32_t jb0 = rowstart[r] + w0;
					for (; s < s1; s += 32u, jb0 += 32u) {
						const float4 *__restrict__ q = stage + s;
						const uint32_t left = s1 - s; // candidates left in the window; groups of 4, at most 8 per mask
						uint32_t mask = 0;
#define CT_TEST(e, bit) do { const float4 q_ = q[e]; \
	const float t_ = __fmaf_rn(nrz, q_.z, __fmaf_rn(nry, q_.y, __fmaf_rn(nrx, q_.x, q_.w))); \
	if (t_ < T) { mask |= (bit); } } while (0)
#pragma unroll
						for (int g = 0; g < 8; ++g) {
							if ((uint32_t)(4 * g) < left) {
								CT_TEST(4 * g + 0, 1u << (4 * g + 0));
								CT_TEST(4 * g + 1, 1u << (4 * g + 1));
								CT_TEST(4 * g + 2, 1u << (4 * g + 2));
								CT_TEST(4 * g + 3, 1u << (4 * g + 3));
							}
						}
#undef CT_TEST
						if (mask) {
							rec[nr < CT_LIST ? nr : CT_LIST - 1] = make_uint2(mask, jb0);
							++nr;
						}
					}
				}
			}
			if (nr > CT_LIST) { // more records than the list holds (very crowded rows): plain fp64 loop instead
				spring_global(G, M, px, py, pz, begin, i, p, sx, sy, sz);
			} else {
				// Phase 2: the recorded candidates in staging order, fp64, from the original positions; the next
				// candidate's position is fetched one iteration ahead of its use
				int k = 0;
				uint32_t m = 0, jb = 0;
				// the record after the current one is fetched while the current one is worked off (the records live in local
				// memory; r3n: 10.7 % of the kernel's stall samples sat on the compare behind that load; r3p: 26.96 -> 26.44 ms)
				uint2 rnext = rec[0];
				auto next_hit = [&](uint32_t &j) -> bool { // advances to the next recorded candidate
					while (m == 0u) {
						if (k >= nr) { return false; }
						m = rnext.x;
						jb = rnext.y;
						++k;
						rnext = rec[k < CT_LIST ? k : CT_LIST - 1];
					}
					j = jb + (uint32_t)(__ffs((int)m) - 1);
					m &= m - 1u;
					return true;
				};
				uint32_t j = 0;
				bool have = next_hit(j);
				double ov[3] = { 0.0, 0.0, 0.0 };
				if (have) { ov[0] = px[j]; ov[1] = py[j]; ov[2] = pz[j]; }
				while (have) {
					uint32_t jn = (uint32_t)i;
					const bool more = next_hit(jn);
					const double on[3] = { px[jn], py[jn], pz[jn] };
					if (j != (uint32_t)i) { pair_exact(M, p, ov, sx, sy, sz); }
					have = more;
					j = jn;
					ov[0] = on[0];
					ov[1] = on[1];
					ov[2] = on[2];
				}
			}
		}
		double np3[3] = { p[0] + sx * M.corr_factor, p[1] + sy * M.corr_factor, p[2] + sz * M.corr_factor };
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			np3[d] = dclamp_std(np3[d], M.gmin[d], M.gmax[d]);
		}
		if (COLLIDE) {
			collide_one(G, M, typ, p, np3);
		}
		nx_[i] = np3[0];
		ny_[i] = np3[1];
		nz_[i] = np3[2];
	}
}

static int correct_impl(lfk_ctx *c, double dt, bool fuse_collide) {
	PhaseTimer T(c, LFK_PHASE_CORRECT_COLLIDE);
	LFK_REQUIRE(c, c->table_valid, LFK_E_STATE, "lfk_correct needs the cell table of lfk_hash");
	if (c->np == 0) { return 0; }
	MotionParams M = motion_params(c, dt);
	const GridDesc &G = c->g;
	dim3 grid((unsigned)((G.nx + CT_LX - 1) / CT_LX), (unsigned)((G.ny + CT_TY - 1) / CT_TY),
		(unsigned)((G.nzl + CT_TZ - 1) / CT_TZ));
	const size_t smem = (size_t)CT_CAP * sizeof(float4);
	static bool attr_set[LFK_MAX_DEVICES] = {}; // function attributes are per device
	if (!attr_set[c->device % LFK_MAX_DEVICES]) {
		LFK_CUDA(c, cudaFuncSetAttribute(k_correct_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		LFK_CUDA(c, cudaFuncSetAttribute(k_correct_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set[c->device % LFK_MAX_DEVICES] = true;
	}
	if (!fuse_collide) { LFK_TRY(materialise_old(c)); }
	if (fuse_collide) {
		LFK_LAUNCH(c, k_correct_tile<true>, grid, CT_THREADS, smem, G, M, c->P, c->Palt.f[PF_PX], c->Palt.f[PF_PY],
			c->Palt.f[PF_PZ], c->begin, c->typ);
	} else {
		LFK_LAUNCH(c, k_correct_tile<false>, grid, CT_THREADS, smem, G, M, c->P, c->Palt.f[PF_PX], c->Palt.f[PF_PY],
			c->Palt.f[PF_PZ], c->begin, c->typ);
	}
	for (int d = 0; d < 3; ++d) {
		double *t = c->P.f[PF_PX + d];
		c->P.f[PF_PX + d] = c->Palt.f[PF_PX + d];
		c->Palt.f[PF_PX + d] = t;
	}
	return 0;
}
int lfkp_correct(lfk_ctx *c, double dt) { return correct_impl(c, dt, false); }
int lfkp_correct_collide(lfk_ctx *c, double dt) {
	if (c->old_valid) {
		LFK_TRY(correct_impl(c, dt, false));
		return lfkp_collide(c);
	}
	return correct_impl(c, dt, true);
}

// =========================================================================================================
// A4: CFL (reference src/simulation.cpp:199-205)
// =========================================================================================================
__global__ void k_max_speed2(const double *__restrict__ vx, const double *__restrict__ vy,
	const double *__restrict__ vz, unsigned long long n, unsigned long long *__restrict__ out_bits) {
	double m = 0.0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
		i += (unsigned long long)gridDim.x * blockDim.x) {
		double s = 0.0;
		s += vx[i] * vx[i];
		s += vy[i] * vy[i];
		s += vz[i] * vz[i];
		m = dmax_std(m, s);
	}
	m = block_max(m);
	if (threadIdx.x == 0) {
		atomicMax(out_bits, (unsigned long long)__double_as_longlong(m)); // non-negative doubles order like u64
	}
}

int lfkp_cfl(lfk_ctx *c, double *value) {
	PhaseTimer T(c, LFK_PHASE_CFL);
	if (c->speed2_valid) { // the last G2P folded the maximum in (g2p.cu); nothing has touched the velocities since
		LFK_CUDA(c, cudaMemcpyAsync(c->d_reduce, c->d_reduce + LFK_REDUCE_SPEED2, sizeof(double), cudaMemcpyDeviceToDevice,
			c->stream));
	} else {
		LFK_CUDA(c, cudaMemsetAsync(c->d_reduce, 0, sizeof(double), c->stream));
	}
	if (c->np > 0 && !c->speed2_valid) {
		unsigned nb = lfk_blocks((long long)c->np, 256);
		if (nb > 148 * 8) { nb = 148 * 8; }
		// (the maximum does not depend on the order, so a velocity payload still in pre-sort order is fine -- but then it
		// lives at the pre-sort indices, hence materialise first when a permutation is pending)
		LFK_TRY(lfkp_materialise_vc(c));
		LFK_LAUNCH(c, k_max_speed2, nb, 256, 0, c->P.f[PF_VX] + c->first, c->P.f[PF_VY] + c->first,
			c->P.f[PF_VZ] + c->first, (unsigned long long)c->np, (unsigned long long*)c->d_reduce);
	}
	if (c->nranks > 1) {
		LFK_TRY(lfkx_allreduce_max(c, c->d_reduce, 1));
	}
	LFK_TRY(lfk_readback(c, c->h_reduce, c->d_reduce, sizeof(double)));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	*value = c->g.h / sqrt(c->h_reduce[0]);
	return 0;
}

// =========================================================================================================
// N1: fluid sources on the device (reference src/simulation.cpp:136-151, 227-238, 756-765)
// =========================================================================================================
// velocity coercion: the reference walks the table of the step's first sort over the source cells; a particle's cell in
// that table is the key of its current position, so the same particles are found by keying every particle
__global__ void k_coerce_sources(GridDesc G, ParticleSoA P, unsigned long long n, const uint16_t *__restrict__ src_map,
	const double *__restrict__ src_vel) {
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	const int x = cell_coord_clamped(P.f[PF_PX][i], G.off[0], G, G.nx);
	const int y = cell_coord_clamped(P.f[PF_PY][i], G.off[1], G, G.ny);
	const int z = cell_coord_clamped(P.f[PF_PZ][i], G.off[2], G, G.nz);
	const int lz = z - G.z0 + 1;
	if (lz < 1 || lz > G.nzl) { return; }
	const unsigned k = src_map[x + (long long)G.nx * (y + (long long)G.ny * lz)];
	if (k == 0) { return; }
	P.f[PF_VX][i] = src_vel[3 * (k - 1)];
	P.f[PF_VY][i] = src_vel[3 * (k - 1) + 1];
	P.f[PF_VZ][i] = src_vel[3 * (k - 1) + 2];
#pragma unroll
	for (int f = PF_C0; f < PF_C0 + 9; ++f) { P.f[f][i] = 0.0; }
}

// seed_cell's bookkeeping: an entry (cell, source) adds max(0, target - num) particles, where num is the cell's count
// after the earlier entries of the same cell (the reference sets _space_hash(cell).count = target after seeding)
__global__ void k_source_need(const uint32_t *__restrict__ src_cell, const uint32_t *__restrict__ src_of,
	const uint32_t *__restrict__ src_target, const uint32_t *__restrict__ cnt, uint32_t *__restrict__ need, uint32_t entries) {
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= entries) { return; }
	const uint32_t cell = src_cell[e];
	uint32_t e0 = e;
	while (e0 > 0 && src_cell[e0 - 1] == cell) { --e0; }
	uint32_t num = cnt[cell], add = 0;
	for (uint32_t k = e0; k <= e; ++k) {
		const uint32_t target = src_target[src_of[k]];
		add = target > num ? target - num : 0u;
		// (the reference assigns count = target even when the cell held more: later entries then see `target`)
		num = target;
	}
	need[e] = add;
}

__global__ void k_source_seed(GridDesc G, ParticleSoA P, uint32_t *__restrict__ key, unsigned long long base,
	const uint32_t *__restrict__ src_cell, const uint32_t *__restrict__ src_gcell, const uint32_t *__restrict__ src_of,
	const double *__restrict__ src_vel, const uint32_t *__restrict__ off, uint32_t entries, unsigned long long seed,
	unsigned long long step, int with_old) {
	const uint32_t e = blockIdx.x;
	if (e >= entries) { return; }
	const uint32_t b = off[e], n = off[e + 1] - b;
	const uint32_t cell = src_cell[e], gcell = src_gcell[e], sidx = src_of[e];
	const int x = (int)(cell % (uint32_t)G.nx), y = (int)((cell / (uint32_t)G.nx) % (uint32_t)G.ny);
	const int z = (int)(cell / (uint32_t)(G.nx * G.ny)) - 1 + G.z0;
	// seed_cell: offset = grid_offset + cell * cell_size, position = offset + U(0, cell_size)^3 (:143-147)
	const double o0 = G.off[0] + (double)x * G.h, o1 = G.off[1] + (double)y * G.h, o2 = G.off[2] + (double)z * G.h;
	for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
		unsigned long long r = mix64(seed ^ mix64(step * 0x9e3779b97f4a7c15ull + gcell) ^ mix64(((unsigned long long)e << 32) | k));
		double j[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			r = mix64(r + 0x9e3779b97f4a7c15ull);
			j[d] = (double)(r >> 11) * (1.0 / 9007199254740992.0) * G.h;
		}
		const unsigned long long i = base + b + k;
		P.f[PF_PX][i] = o0 + j[0];
		P.f[PF_PY][i] = o1 + j[1];
		P.f[PF_PZ][i] = o2 + j[2];
		P.f[PF_VX][i] = src_vel[3 * sidx];
		P.f[PF_VY][i] = src_vel[3 * sidx + 1];
		P.f[PF_VZ][i] = src_vel[3 * sidx + 2];
#pragma unroll
		for (int f = PF_C0; f < PF_C0 + 9; ++f) { P.f[f][i] = 0.0; }
		if (with_old) {
			P.f[PF_OX][i] = o0 + j[0];
			P.f[PF_OY][i] = o1 + j[1];
			P.f[PF_OZ][i] = o2 + j[2];
		}
		key[i] = cell;
	}
}

int lfkp_coerce_sources(lfk_ctx *c) {
	if (!c->src_active || !c->src_coerce || c->np == 0) { return 0; }
	c->speed2_valid = false;
	PhaseTimer T(c, LFK_PHASE_ADVECT_COLLIDE);
	LFK_TRY(lfkp_materialise_vc(c)); // the velocity / c rows are written in particle order
	LFK_LAUNCH(c, k_coerce_sources, lfk_blocks((long long)c->np, 256), 256, 0, c->g, lfk_own_view(c),
		(unsigned long long)c->np, c->src_map, c->src_vel);
	return 0;
}

int lfkp_update_sources(lfk_ctx *c, uint64_t *added) {
	if (added) { *added = 0; }
	if (!c->src_active || c->src_entries == 0) { ++c->rng_step; return 0; }
	c->speed2_valid = false;
	PhaseTimer T(c, LFK_PHASE_SORT);
	LFK_REQUIRE(c, c->table_valid, LFK_E_STATE, "lfk_update_sources needs the cell table of lfk_hash");
	const uint32_t ne = c->src_entries;
	LFK_LAUNCH(c, k_source_need, lfk_blocks(ne, 128), 128, 0, c->src_cell, c->src_of, c->src_target, c->cnt, c->src_need, ne);
	uint32_t *off = c->src_need + (ne + 1);
	LFK_TRY(lfkp_exclusive_scan_u32(c, c->src_need, off, ne, 0));
	uint32_t total = 0;
	LFK_CUDA(c, cudaMemcpyAsync(&total, off + ne, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	const uint64_t step = c->rng_step++;
	if (total == 0) { return 0; }
	LFK_TRY(lfkp_materialise_vc(c)); // appended particles carry their velocity in particle order
	// multi-GPU: the entries behind the own particles are ghost copies of the upper neighbour's boundary layer; the
	// next sort brings fresh ones, so they may be overwritten
	const uint64_t first = c->first, np = c->np;
	LFK_TRY(lfkp_reserve_particles(c, np + total)); // may move the own particles to the front
	LFK_LAUNCH(c, k_source_seed, ne, 64, 0, c->g, c->P, c->key, (unsigned long long)(c->first + np), c->src_cell, c->src_gcell,
		c->src_of, c->src_vel, off, ne, (unsigned long long)c->rng_seed, (unsigned long long)step, c->old_valid ? 1 : 0);
	(void)first;
	c->np = np + total;
	c->ntot = c->first + c->np;
	c->table_valid = false;
	c->ordinal_valid = false;
	c->system_valid = false;
	if (added) { *added = total; }
	return 0;
}

// =========================================================================================================
// Synthetic seeding (bench scenes): jittered sub-cell sampling like simulation::seed_func
// (reference include/fluid/simulation.h:80-115), with a counter-based hash RNG instead of pcg32.
// =========================================================================================================
// candidate `tid` of the seeding lattice: its position, and whether it lies strictly inside the box
struct SeedBox {
	int cx0, cy0, cz0, ex, ey, ez;
	double s0, s1, s2, e0, e1, e2, vx, vy, vz;
	uint32_t dens;
	unsigned long long seed;
};
__device__ __forceinline__ bool seed_candidate(const GridDesc &G, const SeedBox &B, unsigned long long tid,
	double *pos, uint32_t *key) {
	const unsigned long long per_cell = (unsigned long long)B.dens * B.dens * B.dens;
	const unsigned long long total = (unsigned long long)B.ex * B.ey * B.ez * per_cell;
	if (tid >= total) { return false; }
	const unsigned long long sub = tid % per_cell, cell = tid / per_cell;
	const int x = B.cx0 + (int)(cell % B.ex), y = B.cy0 + (int)((cell / B.ex) % B.ey),
		z = B.cz0 + (int)(cell / ((unsigned long long)B.ex * B.ey));
	// reference loop nest: sx outermost, sz innermost
	const int sz_ = (int)(sub % B.dens), sy_ = (int)((sub / B.dens) % B.dens),
		sx_ = (int)(sub / ((unsigned long long)B.dens * B.dens));
	const double small = G.h / (double)B.dens;
	const unsigned long long gid = ((unsigned long long)(x + (long long)G.nx * (y + (long long)G.ny * z))) * per_cell + sub;
	unsigned long long r = mix64(B.seed ^ mix64(gid + 0x9e3779b97f4a7c15ull));
	double j[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		r = mix64(r + 0x9e3779b97f4a7c15ull);
		j[d] = (double)(r >> 11) * (1.0 / 9007199254740992.0) * small;
	}
	pos[0] = G.off[0] + (double)x * G.h + (double)sx_ * small + j[0];
	pos[1] = G.off[1] + (double)y * G.h + (double)sy_ * small + j[1];
	pos[2] = G.off[2] + (double)z * G.h + (double)sz_ * small + j[2];
	*key = (uint32_t)(x + (long long)G.nx * (y + (long long)G.ny * (z - G.z0 + 1)));
	return pos[0] > B.s0 && pos[1] > B.s1 && pos[2] > B.s2 && pos[0] < B.e0 && pos[1] < B.e1 && pos[2] < B.e2;
}
// Two passes so that the particle ORDER is a pure function of the scene (lattice order), not of the scheduling of an
// atomic slot counter: run-to-run and rank-to-rank identical arrays.
#define SEED_THREADS 1024
__global__ void __launch_bounds__(SEED_THREADS) k_seed_count(GridDesc G, SeedBox B, uint32_t *__restrict__ block_cnt) {
	const unsigned long long tid = (unsigned long long)blockIdx.x * SEED_THREADS + threadIdx.x;
	double pos[3];
	uint32_t key;
	const int n = __syncthreads_count(seed_candidate(G, B, tid, pos, &key));
	if (threadIdx.x == 0) { block_cnt[blockIdx.x] = (uint32_t)n; }
}
__global__ void __launch_bounds__(SEED_THREADS) k_seed_write(GridDesc G, SeedBox B, ParticleSoA P,
	uint32_t *__restrict__ keyout, const uint32_t *__restrict__ block_off, unsigned long long base) {
	__shared__ uint32_t warp_tot[32];
	const unsigned long long tid = (unsigned long long)blockIdx.x * SEED_THREADS + threadIdx.x;
	double pos[3];
	uint32_t key;
	const bool in = seed_candidate(G, B, tid, pos, &key);
	const unsigned bal = __ballot_sync(0xffffffffu, in);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (lane == 0) { warp_tot[w] = (uint32_t)__popc(bal); }
	__syncthreads();
	if (!in) { return; }
	uint32_t before = 0;
	for (int k = 0; k < w; ++k) { before += warp_tot[k]; }
	const unsigned long long i = base + block_off[blockIdx.x] + before + (uint32_t)__popc(bal & ((1u << lane) - 1u));
	P.f[PF_PX][i] = pos[0];
	P.f[PF_PY][i] = pos[1];
	P.f[PF_PZ][i] = pos[2];
	P.f[PF_VX][i] = B.vx;
	P.f[PF_VY][i] = B.vy;
	P.f[PF_VZ][i] = B.vz;
#pragma unroll
	for (int f = PF_C0; f < PF_C0 + 9; ++f) { P.f[f][i] = 0.0; }
	keyout[i] = key;
}

int lfkp_seed_box(lfk_ctx *c, const double *start, const double *size, const double *vel, uint32_t dens,
	uint64_t seed, int append) {
	const GridDesc &G = c->g;
	c->speed2_valid = false;
	if (!append) {
		c->np = 0;
		c->first = 0;
		c->ntot = 0;
		c->v_deferred = false;
		c->c_deferred = false;
	}
	LFK_TRY(lfkp_materialise_vc(c));
	double end[3] = { start[0] + size[0], start[1] + size[1], start[2] + size[2] };
	int c0[3], c1[3];
	const int gsz[3] = { G.nx, G.ny, G.nz };
	for (int d = 0; d < 3; ++d) { // world_position_to_cell_index_unclamped + seed_func's clamp of the end
		double a = (start[d] - G.off[d]) / G.h, b = (end[d] - G.off[d]) / G.h;
		long long ca = (long long)(a > 0.0 ? a : 0.0), cb = (long long)(b > 0.0 ? b : 0.0) + 1;
		if (cb > gsz[d]) { cb = gsz[d]; }
		if (ca > cb) { ca = cb; }
		c0[d] = (int)ca;
		c1[d] = (int)cb;
	}
	// this rank seeds only the cells of its slab
	if (c0[2] < G.z0) { c0[2] = G.z0; }
	if (c1[2] > G.z0 + G.nzl) { c1[2] = G.z0 + G.nzl; }
	long long ex = c1[0] - c0[0], ey = c1[1] - c0[1], ez = c1[2] - c0[2];
	if (ex <= 0 || ey <= 0 || ez <= 0) { return 0; }
	unsigned long long per_cell = (unsigned long long)dens * dens * dens;
	unsigned long long total = (unsigned long long)ex * ey * ez * per_cell;
	LFK_TRY(lfkp_reserve_particles(c, c->np + total));
	SeedBox B;
	B.cx0 = c0[0]; B.cy0 = c0[1]; B.cz0 = c0[2];
	B.ex = (int)ex; B.ey = (int)ey; B.ez = (int)ez;
	B.s0 = start[0]; B.s1 = start[1]; B.s2 = start[2];
	B.e0 = end[0]; B.e1 = end[1]; B.e2 = end[2];
	B.vx = vel[0]; B.vy = vel[1]; B.vz = vel[2];
	B.dens = dens;
	B.seed = (unsigned long long)seed;
	const unsigned nb = lfk_blocks((long long)total, SEED_THREADS);
	uint32_t *cnt = nullptr; // [nb + 1] counts, [nb + 1] offsets
	LFK_CUDA(c, cudaMalloc((void**)&cnt, 2 * ((size_t)nb + 1) * sizeof(uint32_t)));
	uint32_t *off = cnt + nb + 1;
	int rc = 0;
	do {
		k_seed_count<<<nb, SEED_THREADS, 0, c->stream>>>(G, B, cnt);
		++c->stats.kernel_launches;
		if ((rc = lfkp_exclusive_scan_u32(c, cnt, off, nb, 0)) != 0) { break; }
		k_seed_write<<<nb, SEED_THREADS, 0, c->stream>>>(G, B, lfk_own_view(c), c->key + c->first, off,
			(unsigned long long)c->np);
		++c->stats.kernel_launches;
		uint32_t h = 0;
		cudaError_t e = cudaMemcpyAsync(&h, off + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) { e = cudaStreamSynchronize(c->stream); }
		if (e == cudaSuccess) { e = cudaGetLastError(); }
		if (e != cudaSuccess) { rc = lfk_fail(c, -(int)e, cudaGetErrorString(e), __FILE__, __LINE__); break; }
		c->np += h;
	} while (0);
	cudaFree(cnt);
	LFK_TRY(rc);
	c->ntot = c->np;

	c->old_valid = false;
	c->table_valid = false;
	c->keys_valid = false;
	return 0;
}
