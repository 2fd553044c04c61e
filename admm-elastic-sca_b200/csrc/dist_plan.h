// dist_plan.h -- host-side plans for ONE mesh partitioned over several ranks (pure host code, no CUDA: also compiled into
// tests/hostcheck, where the plans are checked on the CPU for any world size).
#pragma once
#include <vector>

#include "direct_factor.h"

namespace admmb {

// Subtree-to-rank mapping of the supernodal elimination tree (the classic way to parallelise a sparse triangular solve):
// the tree is cut near its root; everything above the cut (the largest separators) is processed by EVERY rank, each
// subtree below it by ONE rank.  Subtrees only couple through the rows of their common ancestors, so the forward pass
// needs one all-reduce of those (few) rows and the backward pass none.  The cut is chosen by repeatedly opening the
// heaviest subtree and keeping the configuration that minimises (replicated bytes + heaviest rank's bytes), the
// bandwidth cost of the slowest rank.  owner[J] = rank, or -1 for the replicated top.
std::vector<int> shard_owners(const SupernodalFactor &F, int world, double *top_fraction);

// Neighbour-only halo of the row-partitioned system matrix (PCG).  Rank r owns the rows [r0, r1) = its chunk of `chunk`
// consecutive nodes; A_n is structurally symmetric, so "the owned nodes peer q's rows reference" is "the owned rows that
// reference a node of q": both sides derive the same ascending list from their own rows, no negotiation needed.
struct HaloPlan {
	std::vector<int> send_off, send_cnt; // per peer: slice of send_idx
	std::vector<int> send_idx;           // owned nodes (global ids, ascending per peer) whose values the peers need
	std::vector<int> recv_off, recv_cnt; // per peer: slots [recv_off, recv_off + recv_cnt) behind the local vector
	std::vector<std::vector<int> > recv; // per peer: the peer's nodes (global ids, ascending) that land in those slots
	int send_total = 0, recv_total = 0;
	// slot of a column this rank does not own (-1 if no row of this rank references it)
	int slot_of(int column, int chunk) const;
};
void plan_halo(const int *A_ptr, const int *A_idx, int chunk, int world, int r0, int r1, HaloPlan &H);

} // namespace admmb
