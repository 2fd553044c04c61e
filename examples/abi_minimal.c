/* examples/abi_minimal.c -- the C ABI from plain C: the reference's samples/singletet.cpp scene (A/samples/singletet.cpp:
 * 27-111: three anchored nodes, one LinearTetStrain tet of stiffness 1, dt = 1, 20 ADMM iterations, node 4 pulled to
 * x = 200) built and stepped through libadmm_b200.so.  Prints "Node 4 x: 171.571" like the reference.
 *
 *   gcc -std=c99 -Wall -pedantic -I include examples/abi_minimal.c -L admm-elastic-sca_b200 -ladmm_b200 \
 *       -Wl,-rpath,$PWD/admm-elastic-sca_b200 -o abi_minimal && ./abi_minimal
 */
#include <stdio.h>

#include "admm_b200.h"

#define CHECK(call)                                                                 \
	do {                                                                            \
		int rc_ = (call);                                                           \
		if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, admmb_last_error(ctx)); return 1; } \
	} while (0)

int main(void) {
	admmb_ctx *ctx = 0;
	/* node positions and masses, xyz interleaved ("scaled x3", System.hpp:47-49) */
	double x[12] = { 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0 }; /* singletet.cpp:86-91 */
	double m[12] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };
	double v[12] = { 0 };
	const int anchors[3] = { 0, 1, 2 };
	const int tet[4] = { 0, 1, 2, 3 };
	if (admmb_create(0, &ctx) != 0) { fprintf(stderr, "admmb_create: %s\n", admmb_last_error(0)); return 1; }
	CHECK(admmb_set_nodes(ctx, 4, x, m));
	CHECK(admmb_add_static_anchors(ctx, 3, anchors, -1.0));
	CHECK(admmb_add_tets(ctx, ADMMB_TET_LINEAR_STRAIN, 1, tet, 1.0, 0.0, 0.0, 0));
	CHECK(admmb_finalize(ctx, 1.0));
	x[9] = 200.0; /* move node 4 after initialize(), as singletet.cpp:40 does */
	CHECK(admmb_step(ctx, 20, x, v));
	printf("Node 4 x: %g\n", x[9]);
	CHECK(admmb_destroy(ctx));
	return 0;
}
