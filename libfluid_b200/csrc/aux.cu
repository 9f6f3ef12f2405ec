// The two immediate neighbours of the per-step path on SURVEY.md's "next" list, on the device:
//   N2  mesher surface sampling  (reference mesher::_sample_surface_function, src/mesher.cpp:333-376): every frame the
//       testbed / Maya node turn the particle positions into an implicit surface on a finer grid; with the positions
//       resident in HBM that sampling runs here, and only the sampled grid (not 24 B per particle) crosses PCIe;
//   N4  obstacle voxelisation    (reference voxelizer, src/voxelizer.cpp:19-126, and obstacle, src/data_structures/
//       obstacle.cpp:9-29): triangle / box overlap per voxel, flood fill of the exterior, solid cells of the MAC grid.
// Compiled --fmad=false: the voxel classification is integer-valued and must equal the reference's bit for bit, so the
// separating-axis arithmetic keeps the reference's operation order without contraction.
#include "lfk_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

// =========================================================================================================
// N2: mesher surface sampling
// =========================================================================================================
struct MesherDev {
	double off[3], cs, ext2, r;
	int sx, sy, sz, radius;
};

// cell of a particle in the sampling grid: vec3i((p - grid_offset) / cell_size), kept only if every index is > 0 (sic,
// src/mesher.cpp:337-339) and inside the grid (space_hashing::add_object_at)
__global__ void k_mesher_keys(MesherDev M, const double *__restrict__ px, const double *__restrict__ py,
	const double *__restrict__ pz, long long stride, unsigned long long n, uint32_t *__restrict__ key,
	uint32_t *__restrict__ slot, uint32_t *__restrict__ cnt) {
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	const double p[3] = { px[i * stride], py[i * stride], pz[i * stride] };
	int idx[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		const double g = (p[d] - M.off[d]) / M.cs;
		// static_cast<int>: truncation toward zero (values beyond the int range are undefined in the reference)
		idx[d] = g >= 2147483647.0 ? 2147483647 : (g <= -2147483648.0 ? -2147483647 - 1 : (int)g);
	}
	uint32_t k = (uint32_t)((long long)M.sx * M.sy * M.sz); // graveyard
	if (idx[0] > 0 && idx[1] > 0 && idx[2] > 0 && idx[0] < M.sx && idx[1] < M.sy && idx[2] < M.sz) {
		k = (uint32_t)(idx[0] + (long long)M.sx * (idx[1] + (long long)M.sy * idx[2]));
	}
	key[i] = k;
	slot[i] = atomicAdd(cnt + k, 1u);
}
__global__ void k_mesher_scatter(const double *__restrict__ px, const double *__restrict__ py,
	const double *__restrict__ pz, long long stride, unsigned long long n, const uint32_t *__restrict__ key,
	const uint32_t *__restrict__ slot, const uint32_t *__restrict__ begin, double *__restrict__ sx_,
	double *__restrict__ sy_, double *__restrict__ sz_) {
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	const uint32_t at = begin[key[i]] + slot[i];
	sx_[at] = px[i * stride];
	sy_[at] = py[i * stride];
	sz_[at] = pz[i * stride];
}
// one thread per grid point (x fastest); the (2 R)^3 cells around it are walked row by row: a row of cells is one
// contiguous range of the cell-sorted positions
__global__ void __launch_bounds__(128) k_mesher_sample(MesherDev M, const uint32_t *__restrict__ begin,
	const double *__restrict__ qx, const double *__restrict__ qy, const double *__restrict__ qz,
	double *__restrict__ out) {
	const int gx = M.sx + 1, gy = M.sy + 1, gz = M.sz + 1;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)gx * gy * gz) { return; }
	const int x = (int)(t % gx), y = (int)((t / gx) % gy), z = (int)(t / ((long long)gx * gy));
	const double g0 = M.off[0] + M.cs * (double)x, g1 = M.off[1] + M.cs * (double)y, g2 = M.off[2] + M.cs * (double)z;
	const int R = M.radius;
	const int x0 = max(x - R, 0), x1 = min(x + R, M.sx), y0 = max(y - R, 0), y1 = min(y + R, M.sy),
		z0 = max(z - R, 0), z1 = min(z + R, M.sz);
	double tw = 0.0, tr = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0;
	bool has = false;
	if (x0 < x1) {
		for (int cz = z0; cz < z1; ++cz) {
			for (int cy = y0; cy < y1; ++cy) {
				const long long row = (long long)M.sx * (cy + (long long)M.sy * cz);
				const uint32_t b = begin[row + x0], e = begin[row + x1];
				has |= e > b;
				for (uint32_t q = b; q < e; ++q) {
					const double p0 = qx[q], p1 = qy[q], p2 = qz[q];
					const double d0 = p0 - g0, d1 = p1 - g1, d2 = p2 - g2;
					double sq = 0.0; // vec_ops::dot: ((0 + x x) + y y) + z z
					sq += d0 * d0;
					sq += d1 * d1;
					sq += d2 * d2;
					double s = 1.0 - sq / M.ext2; // _kernel(squared_length / (particle_extent * particle_extent))
					double w = 0.0;
					if (s > 0.0) { w = s * s * s; }
					tw += w;
					tr += w * M.r;
					t0 += w * p0;
					t1 += w * p1;
					t2 += w * p2;
				}
			}
		}
	}
	double value = 1.0;
	if (has) { // (a point whose particles all have zero weight divides 0 by 0, like the reference: NaN)
		tr /= tw;
		t0 /= tw;
		t1 /= tw;
		t2 /= tw;
		const double e0 = t0 - g0, e1 = t1 - g1, e2 = t2 - g2;
		double sq = 0.0;
		sq += e0 * e0;
		sq += e1 * e1;
		sq += e2 * e2;
		value = sqrt(sq) - tr;
	}
	out[t] = value;
}

extern "C" int lfk_mesher_sample(lfk_ctx *c, const lfk_mesher *m, double r, const double *xyz, uint64_t n,
	double *surface) {
	if (!c || !m || !surface) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, m->size[0] >= 1 && m->size[1] >= 1 && m->size[2] >= 1 && m->size[0] * m->size[1] * m->size[2] < (1ull << 31),
		LFK_E_INVALID, "mesher grid size out of range");
	LFK_REQUIRE(c, m->cell_size > 0.0 && m->particle_extent > 0.0, LFK_E_INVALID, "mesher cell_size / particle_extent must be > 0");
	MesherDev M;
	for (int d = 0; d < 3; ++d) { M.off[d] = m->grid_offset[d]; }
	M.cs = m->cell_size;
	M.ext2 = m->particle_extent * m->particle_extent;
	M.r = r;
	M.sx = (int)m->size[0]; M.sy = (int)m->size[1]; M.sz = (int)m->size[2];
	M.radius = (int)std::min<uint64_t>(m->cell_radius, 1u << 20);
	const size_t ncell = (size_t)M.sx * M.sy * M.sz, npts = (size_t)(M.sx + 1) * (M.sy + 1) * (M.sz + 1);
	// positions: the caller's (host, xyz triples) or the context's own particles (device SoA)
	const double *px, *py, *pz;
	long long stride = 1;
	double *d_xyz = nullptr;
	if (xyz) {
		LFK_CUDA(c, cudaMalloc((void**)&d_xyz, (size_t)(n ? n : 1) * 24));
		LFK_CUDA(c, cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
		px = d_xyz; py = d_xyz + 1; pz = d_xyz + 2;
		stride = 3;
	} else {
		n = c->np;
		px = c->P.f[PF_PX] + c->first; py = c->P.f[PF_PY] + c->first; pz = c->P.f[PF_PZ] + c->first;
	}
	uint32_t *key = nullptr, *slot = nullptr, *cnt = nullptr, *begin = nullptr;
	double *sorted = nullptr, *d_out = nullptr;
	int rc = 0;
	do {
#define AUX_CUDA(expr) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { rc = lfk_fail(c, -(int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); break; } }
		AUX_CUDA(cudaMalloc((void**)&key, (size_t)(n ? n : 1) * 4));
		AUX_CUDA(cudaMalloc((void**)&slot, (size_t)(n ? n : 1) * 4));
		AUX_CUDA(cudaMalloc((void**)&cnt, (ncell + 2) * 4));
		AUX_CUDA(cudaMalloc((void**)&begin, (ncell + 2) * 4));
		AUX_CUDA(cudaMalloc((void**)&sorted, (size_t)(n ? n : 1) * 24));
		AUX_CUDA(cudaMalloc((void**)&d_out, npts * 8));
		AUX_CUDA(cudaMemsetAsync(cnt, 0, (ncell + 2) * 4, c->stream));
		if (n > 0) {
			k_mesher_keys<<<lfk_blocks((long long)n, 256), 256, 0, c->stream>>>(M, px, py, pz, stride, n, key, slot, cnt);
			++c->stats.kernel_launches;
		}
		if ((rc = lfkp_exclusive_scan_u32(c, cnt, begin, (long long)ncell + 1, 0)) != 0) { break; }
		double *qx = sorted, *qy = sorted + (n ? n : 1), *qz = sorted + 2 * (n ? n : 1);
		if (n > 0) {
			k_mesher_scatter<<<lfk_blocks((long long)n, 256), 256, 0, c->stream>>>(px, py, pz, stride, n, key, slot, begin, qx, qy, qz);
			++c->stats.kernel_launches;
		}
		k_mesher_sample<<<lfk_blocks((long long)npts, 128), 128, 0, c->stream>>>(M, begin, qx, qy, qz, d_out);
		++c->stats.kernel_launches;
		AUX_CUDA(cudaGetLastError());
		AUX_CUDA(cudaMemcpyAsync(surface, d_out, npts * 8, cudaMemcpyDeviceToHost, c->stream));
		AUX_CUDA(cudaStreamSynchronize(c->stream));
	} while (0);
	cudaFree(key); cudaFree(slot); cudaFree(cnt); cudaFree(begin); cudaFree(sorted); cudaFree(d_out); cudaFree(d_xyz);
	return rc;
}

// =========================================================================================================
// N4: voxeliser + obstacle
// =========================================================================================================
struct VoxDev {
	double off[3], cs;
	int sx, sy, sz;
};
enum { VOX_INTERIOR = 0, VOX_EXTERIOR = 1, VOX_SURFACE = 2 }; // voxelizer::cell_type (include/fluid/voxelizer.h:17-21)

// aab_triangle_overlap_bounded (src/math/intersection.cpp:31-82), operation for operation
__device__ bool tri_box_overlap(const double *c, double he, const double *a, const double *b, const double *d) {
	double v[3][3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		v[0][k] = a[k] - c[k];
		v[1][k] = b[k] - c[k];
		v[2][k] = d[k] - c[k];
	}
	double f[3][3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		f[0][k] = v[1][k] - v[0][k];
		f[1][k] = v[2][k] - v[1][k];
		f[2][k] = v[0][k] - v[2][k];
	}
	const double nrm[3] = { f[0][1] * f[1][2] - f[0][2] * f[1][1], f[0][2] * f[1][0] - f[0][0] * f[1][2],
		f[0][0] * f[1][1] - f[0][1] * f[1][0] };
	double center_off = 0.0, radius_n = 0.0;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		center_off += v[0][k] * nrm[k];
		radius_n += fabs(nrm[k]) * he;
	}
	if (fabs(center_off) > fabs(radius_n)) { return false; }
	// nine edge axes: for axis group g (x, y, z) and edge i, the two distinct projections are those of v[i] and v[i + 2]
#pragma unroll
	for (int g = 0; g < 3; ++g) {
		const int p = (g + 2) % 3, q = (g + 1) % 3; // (x: z, y) (y: x, z) (z: y, x)
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			const double *v1 = v[i], *v2 = v[(i + 2) % 3], *fi = f[i];
			const double p0 = v1[p] * fi[q] - v1[q] * fi[p], p1 = v2[p] * fi[q] - v2[q] * fi[p];
			const double pmin = p1 < p0 ? p1 : p0, pmax = p1 < p0 ? p0 : p1; // std::minmax(p0, p1)
			const double r = he * fabs(fi[p]) + he * fabs(fi[q]);
			if (pmin > r || pmax < -r) { return false; }
		}
	}
	return true;
}

// voxelize_triangle (src/voxelizer.cpp:54-79): one block per triangle, threads over the cells of its bounding box;
// the cell centres are built by repeated addition exactly like the reference's loop counters
__global__ void __launch_bounds__(128) k_vox_surface(VoxDev V, const double *__restrict__ pos,
	const unsigned long long *__restrict__ idx, unsigned long long ntri, uint8_t *__restrict__ vox) {
	for (unsigned long long t = blockIdx.x; t < ntri; t += gridDim.x) {
		double a[3], b[3], d[3], lo[3], hi[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			a[k] = pos[3 * idx[3 * t] + k];
			b[k] = pos[3 * idx[3 * t + 1] + k];
			d[k] = pos[3 * idx[3 * t + 2] + k];
			lo[k] = fmin(fmin(a[k], b[k]), d[k]);
			hi[k] = fmax(fmax(a[k], b[k]), d[k]);
		}
		const double half = 0.5 * V.cs;
		unsigned long long mn[3], mx[3];
		double minc[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			mn[k] = (unsigned long long)((lo[k] - V.off[k]) / V.cs);
			mx[k] = (unsigned long long)((hi[k] - V.off[k]) / V.cs);
			minc[k] = V.off[k] + (double)mn[k] * V.cs + half;
		}
		const unsigned long long ex = mx[0] - mn[0] + 1, ey = mx[1] - mn[1] + 1, ez = mx[2] - mn[2] + 1;
		for (unsigned long long e = threadIdx.x; e < ex * ey * ez; e += blockDim.x) {
			const unsigned long long dx = e % ex, dy = (e / ex) % ey, dz = e / (ex * ey);
			double c[3] = { minc[0], minc[1], minc[2] };
			for (unsigned long long k = 0; k < dx; ++k) { c[0] += V.cs; }
			for (unsigned long long k = 0; k < dy; ++k) { c[1] += V.cs; }
			for (unsigned long long k = 0; k < dz; ++k) { c[2] += V.cs; }
			const unsigned long long x = mn[0] + dx, y = mn[1] + dy, z = mn[2] + dz;
			if (x >= (unsigned long long)V.sx || y >= (unsigned long long)V.sy || z >= (unsigned long long)V.sz) { continue; }
			uint8_t *cell = vox + (x + (unsigned long long)V.sx * (y + (unsigned long long)V.sy * z));
			if (*cell != VOX_SURFACE && tri_box_overlap(c, half, a, b, d)) { *cell = VOX_SURFACE; }
		}
	}
}

// mark_exterior (src/voxelizer.cpp:81-126): the cells reachable from (0, 0, 0) through interior cells.  The reference
// walks them with a stack; the reachable SET does not depend on the order, so it is grown here by sweeps until nothing
// changes.  A block relaxes its 8^3 brick (plus halo) in shared memory until stable before writing back, which cuts the
// number of global sweeps from the grid's diameter in cells to its diameter in bricks.
#define VB 8
__global__ void __launch_bounds__(512) k_vox_flood(VoxDev V, uint8_t *__restrict__ vox, int *__restrict__ changed) {
	__shared__ uint8_t tile[VB + 2][VB + 2][VB + 2];
	__shared__ int again;
	const int bx = blockIdx.x * VB, by = blockIdx.y * VB, bz = blockIdx.z * VB;
	for (int e = threadIdx.x; e < (VB + 2) * (VB + 2) * (VB + 2); e += blockDim.x) {
		const int lx = e % (VB + 2), ly = (e / (VB + 2)) % (VB + 2), lz = e / ((VB + 2) * (VB + 2));
		const int x = bx + lx - 1, y = by + ly - 1, z = bz + lz - 1;
		uint8_t v = VOX_SURFACE; // outside the grid: a wall for the flood
		if (x >= 0 && y >= 0 && z >= 0 && x < V.sx && y < V.sy && z < V.sz) {
			v = vox[x + (long long)V.sx * (y + (long long)V.sy * z)];
		}
		tile[lz][ly][lx] = v;
	}
	__syncthreads();
	const int lx = threadIdx.x % VB + 1, ly = (threadIdx.x / VB) % VB + 1, lz = threadIdx.x / (VB * VB) + 1;
	bool mine_changed = false;
	for (int it = 0; it < 3 * VB; ++it) {
		if (threadIdx.x == 0) { again = 0; }
		__syncthreads();
		if (tile[lz][ly][lx] == VOX_INTERIOR &&
			(tile[lz][ly][lx - 1] == VOX_EXTERIOR || tile[lz][ly][lx + 1] == VOX_EXTERIOR ||
			 tile[lz][ly - 1][lx] == VOX_EXTERIOR || tile[lz][ly + 1][lx] == VOX_EXTERIOR ||
			 tile[lz - 1][ly][lx] == VOX_EXTERIOR || tile[lz + 1][ly][lx] == VOX_EXTERIOR)) {
			tile[lz][ly][lx] = VOX_EXTERIOR; // (a neighbour read in the same sweep sees the old or the new value: both fine)
			mine_changed = true;
			again = 1;
		}
		__syncthreads();
		if (!again) { break; }
		__syncthreads();
	}
	if (mine_changed) {
		const int x = bx + lx - 1, y = by + ly - 1, z = bz + lz - 1;
		if (x < V.sx && y < V.sy && z < V.sz) {
			vox[x + (long long)V.sx * (y + (long long)V.sy * z)] = VOX_EXTERIOR;
			*changed = 1;
		}
	}
}
__global__ void k_vox_seed(uint8_t *vox) {
	if (vox[0] != VOX_SURFACE) { vox[0] = VOX_EXTERIOR; }
}

// obstacle cells: interior voxels that lie inside the simulation grid, in the reference's order (z, y, x ascending)
__global__ void k_obstacle_flags(VoxDev V, const uint8_t *__restrict__ vox, long long gx, long long gy, long long gz,
	int nx, int ny, int nz, uint32_t *__restrict__ flag) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long long)V.sx * V.sy * V.sz) { return; }
	const long long x = i % V.sx + gx, y = (i / V.sx) % V.sy + gy, z = i / ((long long)V.sx * V.sy) + gz;
	flag[i] = (vox[i] == VOX_INTERIOR && x >= 0 && y >= 0 && z >= 0 && x < nx && y < ny && z < nz) ? 1u : 0u;
}
__global__ void k_obstacle_emit(VoxDev V, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ at,
	long long gx, long long gy, long long gz, unsigned long long *__restrict__ cells, GridDesc G,
	uint8_t *__restrict__ typ, int mark_solid) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (long long)V.sx * V.sy * V.sz || !flag[i]) { return; }
	const long long x = i % V.sx + gx, y = (i / V.sx) % V.sy + gy, z = i / ((long long)V.sx * V.sy) + gz;
	if (cells) {
		cells[3 * (size_t)at[i]] = (unsigned long long)x;
		cells[3 * (size_t)at[i] + 1] = (unsigned long long)y;
		cells[3 * (size_t)at[i] + 2] = (unsigned long long)z;
	}
	if (mark_solid) {
		const long long lz = z - G.z0 + 1;
		if (lz >= 0 && lz < G.nlz) { typ[x + (long long)G.nx * (y + (long long)G.ny * lz)] = LFK_CELL_SOLID; }
	}
}

extern "C" int lfk_voxelize_mesh(lfk_ctx *c, const double *positions, uint64_t nverts, const uint64_t *indices,
	uint64_t nidx, double cell_size, const double ref_offset[3], int64_t grid_min[3], uint64_t vox_size[3]) {
	if (!c || !positions || !indices || !ref_offset || !grid_min || !vox_size) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, nverts > 0 && nidx >= 3 && cell_size > 0.0, LFK_E_INVALID, "empty mesh or bad cell size");
	for (uint64_t i = 0; i < nidx; ++i) { LFK_REQUIRE(c, indices[i] < nverts, LFK_E_INVALID, "mesh index out of range"); }
	// voxelizer::get_bounding_box + resize_reposition_grid_constrained (include/fluid/voxelizer.h:24-36,
	// src/voxelizer.cpp:19-35), on the host like the reference (a few flops per vertex)
	double mn[3], mx[3];
	for (int d = 0; d < 3; ++d) { mn[d] = mx[d] = positions[d]; }
	for (uint64_t i = 1; i < nverts; ++i) {
		for (int d = 0; d < 3; ++d) {
			mn[d] = std::min(mn[d], positions[3 * i + d]);
			mx[d] = std::max(mx[d], positions[3 * i + d]);
		}
	}
	VoxDev V;
	V.cs = cell_size;
	long long gmin[3], gmax[3];
	for (int d = 0; d < 3; ++d) {
		gmin[d] = (long long)(int)std::floor((mn[d] - ref_offset[d]) / cell_size) - 1;
		gmax[d] = (long long)(int)std::ceil((mx[d] - ref_offset[d]) / cell_size) + 1;
		V.off[d] = ref_offset[d] + (double)gmin[d] * cell_size;
		grid_min[d] = gmin[d];
		vox_size[d] = (uint64_t)(gmax[d] - gmin[d]);
	}
	LFK_REQUIRE(c, vox_size[0] * vox_size[1] * vox_size[2] < (1ull << 31), LFK_E_INVALID, "voxel grid too large");
	V.sx = (int)vox_size[0]; V.sy = (int)vox_size[1]; V.sz = (int)vox_size[2];
	const size_t nvox = (size_t)V.sx * V.sy * V.sz;
	if (c->vox) { cudaFree(c->vox); c->vox = nullptr; }
	LFK_CUDA(c, cudaMalloc((void**)&c->vox, nvox ? nvox : 1));
	for (int d = 0; d < 3; ++d) { c->vox_min[d] = gmin[d]; c->vox_size[d] = (int)vox_size[d]; }
	LFK_CUDA(c, cudaMemsetAsync(c->vox, VOX_INTERIOR, nvox ? nvox : 1, c->stream));
	double *d_pos = nullptr;
	unsigned long long *d_idx = nullptr;
	int *d_changed = nullptr;
	int rc = 0;
	do {
		AUX_CUDA(cudaMalloc((void**)&d_pos, (size_t)nverts * 24));
		AUX_CUDA(cudaMalloc((void**)&d_idx, (size_t)nidx * 8));
		AUX_CUDA(cudaMalloc((void**)&d_changed, sizeof(int)));
		AUX_CUDA(cudaMemcpyAsync(d_pos, positions, (size_t)nverts * 24, cudaMemcpyHostToDevice, c->stream));
		AUX_CUDA(cudaMemcpyAsync(d_idx, indices, (size_t)nidx * 8, cudaMemcpyHostToDevice, c->stream));
		const unsigned long long ntri = nidx / 3; // (i + 2 < indices.size(), include/fluid/voxelizer.h:58)
		k_vox_surface<<<(unsigned)std::min<unsigned long long>(ntri, 65535), 128, 0, c->stream>>>(V, d_pos, d_idx, ntri, c->vox);
		++c->stats.kernel_launches;
		k_vox_seed<<<1, 1, 0, c->stream>>>(c->vox);
		++c->stats.kernel_launches;
		dim3 grid((unsigned)((V.sx + VB - 1) / VB), (unsigned)((V.sy + VB - 1) / VB), (unsigned)((V.sz + VB - 1) / VB));
		for (int sweep = 0; sweep < 100000; ++sweep) {
			AUX_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), c->stream));
			k_vox_flood<<<grid, VB * VB * VB, 0, c->stream>>>(V, c->vox, d_changed);
			++c->stats.kernel_launches;
			int h = 0;
			AUX_CUDA(cudaMemcpyAsync(&h, d_changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
			AUX_CUDA(cudaStreamSynchronize(c->stream));
			if (!h) { break; }
		}
		if (rc) { break; }
		AUX_CUDA(cudaGetLastError());
	} while (0);
	cudaFree(d_pos); cudaFree(d_idx); cudaFree(d_changed);
	return rc;
}

extern "C" int lfk_voxels_download(lfk_ctx *c, uint8_t *voxels, uint64_t capacity) {
	if (!c || !voxels) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->vox != nullptr, LFK_E_STATE, "no voxel grid (call lfk_voxelize_mesh)");
	const size_t nvox = (size_t)c->vox_size[0] * c->vox_size[1] * c->vox_size[2];
	LFK_REQUIRE(c, capacity >= nvox, LFK_E_CAPACITY, "voxel buffer too small");
	LFK_CUDA(c, cudaMemcpyAsync(voxels, c->vox, nvox, cudaMemcpyDeviceToHost, c->stream));
	LFK_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int lfk_obstacle_cells(lfk_ctx *c, uint64_t *cells_xyz, uint64_t capacity, uint64_t *n, int mark_solid) {
	if (!c || !n) { return LFK_E_INVALID; }
	LFK_REQUIRE(c, c->vox != nullptr, LFK_E_STATE, "no voxel grid (call lfk_voxelize_mesh)");
	VoxDev V{};
	V.sx = c->vox_size[0]; V.sy = c->vox_size[1]; V.sz = c->vox_size[2];
	const size_t nvox = (size_t)V.sx * V.sy * V.sz;
	uint32_t *flag = nullptr, *at = nullptr;
	unsigned long long *d_cells = nullptr;
	int rc = 0;
	do {
		AUX_CUDA(cudaMalloc((void**)&flag, (nvox + 2) * 4));
		AUX_CUDA(cudaMalloc((void**)&at, (nvox + 2) * 4));
		k_obstacle_flags<<<lfk_blocks((long long)nvox, 256), 256, 0, c->stream>>>(V, c->vox, c->vox_min[0], c->vox_min[1],
			c->vox_min[2], c->g.nx, c->g.ny, c->g.nz, flag);
		++c->stats.kernel_launches;
		if ((rc = lfkp_exclusive_scan_u32(c, flag, at, (long long)nvox, 0)) != 0) { break; }
		uint32_t total = 0;
		AUX_CUDA(cudaMemcpyAsync(&total, at + nvox, 4, cudaMemcpyDeviceToHost, c->stream));
		AUX_CUDA(cudaStreamSynchronize(c->stream));
		*n = total;
		const bool want = cells_xyz != nullptr;
		if (want && capacity < total) { rc = lfk_fail(c, LFK_E_CAPACITY, "obstacle cell buffer too small", __FILE__, __LINE__); break; }
		if (want && total) { AUX_CUDA(cudaMalloc((void**)&d_cells, (size_t)total * 24)); }
		if (total && (want || mark_solid)) {
			k_obstacle_emit<<<lfk_blocks((long long)nvox, 256), 256, 0, c->stream>>>(V, flag, at, c->vox_min[0], c->vox_min[1],
				c->vox_min[2], d_cells, c->g, c->typ, mark_solid ? 1 : 0);
			++c->stats.kernel_launches;
			if (want) { AUX_CUDA(cudaMemcpyAsync(cells_xyz, d_cells, (size_t)total * 24, cudaMemcpyDeviceToHost, c->stream)); }
			AUX_CUDA(cudaStreamSynchronize(c->stream));
			if (mark_solid) { c->system_valid = false; }
		}
	} while (0);
	cudaFree(flag); cudaFree(at); cudaFree(d_cells);
	return rc;
}
