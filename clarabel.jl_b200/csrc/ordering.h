// Host-side fill-reducing orderings (see ordering.cpp).
#pragma once
#include <cstdint>
#include <vector>

namespace cb200 {
void build_sym_graph(int64_t n, const int64_t* Ap, const int64_t* Ai,
                     std::vector<int64_t>& xadj, std::vector<int32_t>& adj);
void amd_order_graph(int32_t n, const int64_t* xadj, const int32_t* adj, double dense_scale,
                     int32_t* perm_out);
void nd_order_graph(int32_t n, const int64_t* xadj, const int32_t* adj, double dense_scale,
                    int32_t leaf_size, int32_t* perm_out);
// returns false (perm_out untouched) only when skip_unsplit is set and no separator was found at the
// top level, i.e. the ordering would have been amd_order_graph of the whole graph
bool nd_order_graph_blocks(int32_t n, const int64_t* xadj, const int32_t* adj, double dense_scale,
                           int32_t leaf_size, const int32_t* block_id, int32_t* perm_out,
                           bool skip_unsplit = false);
// PSD-aware ordering (see ordering.cpp): vertices outside the dense cone blocks first (AMD-class),
// then the blocks in a nested-dissection order of the block quotient graph.  false = not applicable.
bool order_blocks_last_nd(int32_t n, const int64_t* xadj, const int32_t* adj, double dense_scale,
                          const int32_t* block_id, int32_t min_blocks, int32_t* perm_out);
}  // namespace cb200

extern "C" {
int32_t cb200_order_amd(int64_t n, const int64_t* colptr, const int64_t* rowval,
                        double dense_scale, int64_t* perm);
int32_t cb200_order_nd(int64_t n, const int64_t* colptr, const int64_t* rowval,
                       double dense_scale, int64_t leaf_size, int64_t* perm);
}
