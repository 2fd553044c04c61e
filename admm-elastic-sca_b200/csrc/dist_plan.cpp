// dist_plan.cpp -- see dist_plan.h
#include "dist_plan.h"

#include <algorithm>

namespace admmb {

std::vector<int> shard_owners(const SupernodalFactor &F, int world, double *top_fraction) {
	const int nb = F.nb;
	std::vector<double> wgt(nb), sub(nb, 0.0);
	std::vector<std::vector<int> > kids(nb);
	double total = 0.0;
	for (int J = 0; J < nb; ++J) {
		const double w = F.start[J + 1] - F.start[J], r = F.rptr[J + 1] - F.rptr[J];
		wgt[J] = (w + r) * w;
		total += wgt[J];
	}
	for (int J = 0; J < nb; ++J) { // children precede their parents in elimination order
		sub[J] += wgt[J];
		if (F.parent[J] >= 0) { sub[F.parent[J]] += sub[J]; kids[F.parent[J]].push_back(J); }
	}
	std::vector<char> is_top(nb, 0);
	std::vector<int> cand;
	for (int J = 0; J < nb; ++J) if (F.parent[J] < 0) cand.push_back(J);
	auto lpt = [&](const std::vector<int> &c, std::vector<int> *assign) {
		std::vector<int> order(c);
		std::sort(order.begin(), order.end(), [&](int a, int b) { return sub[a] > sub[b] || (sub[a] == sub[b] && a < b); });
		std::vector<double> load(world, 0.0);
		if (assign) assign->assign(nb, -2);
		for (int J : order) {
			int r = 0;
			for (int q = 1; q < world; ++q) if (load[q] < load[r]) r = q;
			load[r] += sub[J];
			if (assign) (*assign)[J] = r;
		}
		return *std::max_element(load.begin(), load.end());
	};
	double top_w = 0.0, best = 1e300;
	std::vector<char> best_top;
	std::vector<int> best_cand;
	for (int it = 0; it < 64 * world; ++it) {
		const double cost = top_w + lpt(cand, nullptr);
		if ((int)cand.size() >= world && cost < best) { best = cost; best_top = is_top; best_cand = cand; }
		int h = -1;
		for (size_t i = 0; i < cand.size(); ++i) if (!kids[cand[i]].empty() && (h < 0 || sub[cand[i]] > sub[cand[h]])) h = (int)i;
		if (h < 0) break;
		const int J = cand[h];
		is_top[J] = 1;
		top_w += wgt[J];
		cand.erase(cand.begin() + h);
		cand.insert(cand.end(), kids[J].begin(), kids[J].end());
	}
	std::vector<int> owner(nb, -1);
	if (best_cand.empty()) { if (top_fraction) *top_fraction = 1.0; return owner; } // cannot be cut: everything replicated
	std::vector<int> assign;
	lpt(best_cand, &assign);
	for (int J = nb - 1; J >= 0; --J) { // parents before children
		if (best_top[J]) owner[J] = -1;
		else if (assign[J] >= 0) owner[J] = assign[J];
		else owner[J] = (F.parent[J] >= 0) ? owner[F.parent[J]] : 0;
	}
	double tw = 0.0;
	for (int J = 0; J < nb; ++J) if (owner[J] < 0) tw += wgt[J];
	if (top_fraction) *top_fraction = total > 0.0 ? tw / total : 1.0;
	return owner;
}

int HaloPlan::slot_of(int column, int chunk) const {
	const int o = column / chunk;
	if (o < 0 || o >= (int)recv.size()) return -1;
	const std::vector<int> &L = recv[o];
	const std::vector<int>::const_iterator it = std::lower_bound(L.begin(), L.end(), column);
	if (it == L.end() || *it != column) return -1;
	return recv_off[o] + (int)(it - L.begin());
}

void plan_halo(const int *A_ptr, const int *A_idx, int chunk, int world, int r0, int r1, HaloPlan &H) {
	std::vector<std::vector<int> > send(world);
	H.recv.assign(world, std::vector<int>());
	for (int i = r0; i < r1; ++i)
		for (int q = A_ptr[i]; q < A_ptr[i + 1]; ++q) {
			const int j = A_idx[q];
			if (j >= r0 && j < r1) continue;
			const int o = j / chunk;
			H.recv[o].push_back(j);
			if (send[o].empty() || send[o].back() != i) send[o].push_back(i);
		}
	H.send_off.assign(world, 0); H.send_cnt.assign(world, 0); H.recv_off.assign(world, 0); H.recv_cnt.assign(world, 0);
	H.send_idx.clear();
	H.send_total = H.recv_total = 0;
	for (int o = 0; o < world; ++o) {
		std::sort(H.recv[o].begin(), H.recv[o].end());
		H.recv[o].erase(std::unique(H.recv[o].begin(), H.recv[o].end()), H.recv[o].end());
		H.recv_off[o] = H.recv_total; H.recv_cnt[o] = (int)H.recv[o].size(); H.recv_total += H.recv_cnt[o];
		H.send_off[o] = H.send_total; H.send_cnt[o] = (int)send[o].size(); H.send_total += H.send_cnt[o];
		H.send_idx.insert(H.send_idx.end(), send[o].begin(), send[o].end());
	}
}

} // namespace admmb
