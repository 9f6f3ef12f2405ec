// Multigrid preconditioner (placeholder: Jacobi until the V-cycle lands).
#include "lfk_internal.cuh"

int lfkm_setup(lfk_ctx *c, double a_scale) {
	(void)a_scale;
	c->mg_valid = true;
	return 0;
}
__global__ void k_mg_fallback_jacobi(GridDesc G, const uint8_t *__restrict__ flags, const double *__restrict__ r,
	double *__restrict__ z, double a_scale, const PcgScalars *scal) {
	if (scal->done) { return; }
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long c = own + G.sxy;
	unsigned f = flags[c];
	double out = 0.0;
	if ((f & 8u) && (f & 7u) > 0) { out = r[c] / (a_scale * (double)(f & 7u)); }
	z[c] = out;
}
int lfkm_apply(lfk_ctx *c, const double *r, double *z, double a_scale) {
	LFK_LAUNCH(c, k_mg_fallback_jacobi, lfk_blocks(c->g.nown, 256), 256, 0, c->g, c->flags, r, z, a_scale, c->d_scal);
	return 0;
}
int lfkm_free(lfk_ctx *c) {
	(void)c;
	return 0;
}
