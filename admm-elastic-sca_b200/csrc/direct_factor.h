// direct_factor.h -- host supernodal Cholesky of the scalar system matrix A_n (setup path; the reference does
// this step with Eigen::SimplicialLDLT in System::initialize(), System.cpp:138-140).
#pragma once
#include <string>
#include <vector>

namespace admmb {

// A = L L^T on the nested-dissection supernode partition.  For every supernode J with columns
// [start[J], start[J+1]) and below-diagonal row structure rows[rptr[J]..rptr[J+1]) the factor is kept in the
// "inverse multifrontal" form the device solve consumes:
//     T_J = [ inv(L_JJ) ;  L_RJ * inv(L_JJ) ]        ((w + |R|) x w, column-major, ld = w + |R|)
// so that both triangular solves become independent dense panel products per level of the supernodal
// elimination tree (see direct_solve.cu).
struct SupernodalFactor {
	int n = 0, nb = 0, nlevels = 0;
	std::vector<int> start;     // nb+1
	std::vector<int> rptr;      // nb+1
	std::vector<int> rows;      // concatenated row structures (internal node ids, ascending)
	std::vector<int> parent;    // supernodal elimination tree (-1 for roots)
	std::vector<int> level;     // 0 = leaves; level[parent] > level[child]
	std::vector<size_t> toff;   // nb+1 offsets into T
	std::vector<double> T;
	long nnz_L = 0;             // scalar nonzeros of L incl. diagonal, as stored (dense supernode panels)
	double seconds_symbolic = 0, seconds_numeric = 0;
	int device_fronts = 0;      // fronts whose dense work ran on the device (FrontBackend)
};

// Optional accelerator for the dense work of the large fronts near the root of the elimination tree (where almost
// all of the flops are): the host assembles the front, the backend factors it.  front_gpu.cu implements it on the
// device (cuSOLVER potrf + cuBLAS trsm / syrk / trmm, FP64); without a backend everything runs on the host.
struct FrontBackend {
	virtual ~FrontBackend() {}
	virtual int min_front() const = 0;                 // fronts with w + |R| >= min_front go to the backend
	virtual void reserve(size_t front_doubles, size_t panel_doubles) = 0; // largest m*m and m*w that will be asked for
	virtual double *front_buffer(size_t doubles) = 0;  // host buffer to assemble the front in (page-locked); nullptr = unavailable
	// Fr: assembled front, lower triangle, column-major, ld = m = w + r.  Writes T (m x w, ld = m) =
	// [inv(L11); L21 inv(L11)] and, if U != nullptr, the Schur complement F22 - L21 L21^T (r x r, ld = r, lower part).
	// Returns 0, 1 = not positive definite, -1 = backend failure (the caller then factors this front on the host).
	virtual int factor_front(int m, int w, const double *Fr, double *T, double *U) = 0;
};

// Ap/Ai/Ax: CSR of the full symmetric matrix with sorted columns (internal order).  block_end: end offsets of
// the dissection blocks in elimination order.  Returns 0, or -1 with `err` set (matrix not positive definite).
int supernodal_factorize(int n, const int *Ap, const int *Ai, const double *Ax, const std::vector<int> &block_end,
                         SupernodalFactor &F, std::string &err, FrontBackend *backend = nullptr);

} // namespace admmb
