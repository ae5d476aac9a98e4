// bh_export.h — host-side expansion of the implicit preorder cell array into the
// reference's full visitQuads order (BarnesHutAlg.kt:265-274): every cell incl. the empty
// leaves that subdivide() allocates (BH.kt:159-166).  Debug overlay / parity artefact
// only — not on the per-step path.  Shared by bh_engine.cu (after a device->host copy)
// and the CPU emulation in tests/emul/.
#ifndef BH_EXPORT_H
#define BH_EXPORT_H

#include <stdint.h>
#include <vector>
#include "bh_core.h"

struct BhHostTree {
    BhRoot root;
    int n_in, M;
    const uint64_t* keys;     // [n_in] sorted
    const int* order;         // [n_in] body index per sorted position
    const int* S;             // [n_in+1]
    const BhCellS* sk;        // [M] skeletons
    const BhCellD* cd;        // [M] exact f64 records
    const int* jflag;         // [n_in] jitter replay flags per sorted position, or nullptr
};

struct BhCellsOut {
    int64_t count = 0;
    int64_t cap = 0;
    double *cx = nullptr, *cy = nullptr, *h = nullptr, *mass = nullptr, *comx = nullptr, *comy = nullptr;
    int32_t* body = nullptr;
    void put(double ccx, double ccy, double hh, double m, double x, double y, int32_t b) {
        if (count < cap) {
            if (cx) cx[count] = ccx;
            if (cy) cy[count] = ccy;
            if (h) h[count] = hh;
            if (mass) mass[count] = m;
            if (comx) comx[count] = x;
            if (comy) comy[count] = y;
            if (body) body[count] = b;
        }
        ++count;
    }
};

namespace bh_export_detail {
struct Ctx {
    const BhHostTree& t;
    BhCellsOut& out;
    std::vector<int> first;         // sorted index of the leftmost body below each entry
    std::vector<unsigned char> leaf;
};

inline void rec(Ctx& c, int p, double cx, double cy, double h) {
    const BhHostTree& t = c.t;
    const int d = t.sk[p].level;
    if (c.leaf[p]) {
        c.out.put(cx, cy, h, t.cd[p].mass, t.cd[p].comx, t.cd[p].comy, (int32_t)t.order[c.first[p]]);
        return;
    }
    c.out.put(cx, cy, h, t.cd[p].mass, t.cd[p].comx, t.cd[p].comy, -2);
    const double hh = h / 2.0;
    int ch = p + 1;
    const int end = t.sk[p].skip;
    if (d >= t.root.levels) {
        // jitter-regime cluster (bh_jitter_cluster): C's surviving bodies sit in the child their
        // (mutated) position selects; a "dead" child is an internal cell with four empty leaves;
        // dropped bodies (ghost leaves) are not part of the reference's tree
        const int mask = t.jflag ? (t.jflag[c.first[ch]] >> 4) & 15 : 0;
        for (int dig = 0; dig < 4; ++dig) {
            const double ccx = (dig & 1) ? cx + hh : cx - hh;
            const double ccy = (dig & 2) ? cy + hh : cy - hh;
            int found = -1;
            for (int q = ch; q < end; q = t.sk[q].skip) {
                if (t.jflag && (t.jflag[c.first[q]] & 1)) continue;
                const int qd = (t.cd[q].comx < cx ? 0 : 1) + (t.cd[q].comy < cy ? 0 : 2);
                if (qd == dig) { found = q; break; }
            }
            if (found >= 0) {
                c.out.put(ccx, ccy, hh, t.cd[found].mass, t.cd[found].comx, t.cd[found].comy, (int32_t)t.order[c.first[found]]);
            } else if (mask & (1 << dig)) {
                const double qh = hh / 2.0;
                c.out.put(ccx, ccy, hh, 0.0, ccx, ccy, -2);
                for (int g = 0; g < 4; ++g) {
                    const double gx = (g & 1) ? ccx + qh : ccx - qh, gy = (g & 2) ? ccy + qh : ccy - qh;
                    c.out.put(gx, gy, qh, 0.0, gx, gy, -1);
                }
            } else {
                c.out.put(ccx, ccy, hh, 0.0, ccx, ccy, -1);
            }
        }
        return;
    }
    const int sh = 2 * (t.root.levels - 1 - d);
    for (int dig = 0; dig < 4; ++dig) {
        const double ccx = (dig & 1) ? cx + hh : cx - hh;   // Quad.child, BH.kt:73-80
        const double ccy = (dig & 2) ? cy + hh : cy - hh;
        if (ch < end && (int)((t.keys[c.first[ch]] >> sh) & 3ull) == dig) {
            rec(c, ch, ccx, ccy, hh);
            ch = t.sk[ch].skip;
        } else {
            c.out.put(ccx, ccy, hh, 0.0, ccx, ccy, -1);      // empty leaf, BH.kt:179-183
        }
    }
}
}  // namespace bh_export_detail

inline void bh_export_cells(const BhHostTree& t, BhCellsOut& out) {
    if (t.M == 0) {   // no body in the box: the tree is one empty root leaf
        out.put(t.root.cx, t.root.cy, t.root.half, 0.0, t.root.cx, t.root.cy, -1);
        return;
    }
    bh_export_detail::Ctx c{t, out, std::vector<int>((size_t)t.M, 0), std::vector<unsigned char>((size_t)t.M, 0)};
    for (int i = 0; i < t.n_in; ++i) {
        const int lp = t.S[i + 1] + i;
        for (int p = t.S[i] + i; p < lp; ++p) c.first[p] = i;
        c.first[lp] = i;
        c.leaf[lp] = 1;
    }
    bh_export_detail::rec(c, 0, t.root.cx, t.root.cy, t.root.half);
}

#endif  // BH_EXPORT_H
