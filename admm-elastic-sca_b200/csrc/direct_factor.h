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
};

// Ap/Ai/Ax: CSR of the full symmetric matrix with sorted columns (internal order).  block_end: end offsets of
// the dissection blocks in elimination order.  Returns 0, or -1 with `err` set (matrix not positive definite).
int supernodal_factorize(int n, const int *Ap, const int *Ai, const double *Ax, const std::vector<int> &block_end,
                         SupernodalFactor &F, std::string &err);

} // namespace admmb
