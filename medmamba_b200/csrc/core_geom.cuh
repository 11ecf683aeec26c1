// Block geometry of the fused SS2D core kernels (forward and backward share it).
#pragma once
#include <stdint.h>

namespace mmb {

constexpr int kTrainCap = 8;   // steps per block when state checkpoints are written (= backward block)

// A direction's sequence is cut into blocks of at most `cap` consecutive steps:
//   row view  (k = 0, 2): NB_row blocks of T_row consecutive positions;
//   column view (k = 1, 3): NO_col column groups (nw columns each) x NI_col row blocks (T_col rows each);
//                           either nw == 1 or T_col == H.
struct CoreGeom {
    int T_row, NB_row, nw, T_col, NI_col, NO_col, cap;
    int nblocks_max() const { const int c = NO_col * NI_col; return NB_row > c ? NB_row : c; }
};

inline bool core_geometry(int H, int W, int cap, CoreGeom& g) {
    const int L = H * W;
    g.NB_row = (L + cap - 1) / cap;
    g.T_row = (L + g.NB_row - 1) / g.NB_row;
    if (H <= cap) {
        const int maxw = cap / H;
        g.NO_col = (W + maxw - 1) / maxw;
        g.nw = (W + g.NO_col - 1) / g.NO_col;
        g.T_col = H; g.NI_col = 1;
    } else {
        g.nw = 1; g.NO_col = W;
        g.NI_col = (H + cap - 1) / cap;
        g.T_col = (H + g.NI_col - 1) / g.NI_col;
    }
    if (g.T_row > 256 || g.T_col > 256 || g.nw > 256) return false;
    g.cap = g.T_row > g.nw * g.T_col ? g.T_row : g.nw * g.T_col;
    return true;
}

}  // namespace mmb
