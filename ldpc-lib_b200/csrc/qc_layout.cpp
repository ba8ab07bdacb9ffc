// Host-side construction of the QC edge lists (see qc_layout.h).
#include "qc_layout.h"

namespace ldpcb200 {

bool QcHost::build(const int16_t* hd_, int b_, int c_, int Z_)
{
    if (!hd_ || b_ <= 0 || c_ <= 0 || Z_ <= 0 || c_ > 255 || Z_ > 4095) return false;
    b = b_; c = c_; Z = Z_; N = c * Z; R = b * Z;
    hd.assign(hd_, hd_ + (size_t)b * c);
    rp.assign(b + 1, 0); cp.assign(c + 1, 0);
    col.clear(); sh.clear(); row.clear(); cedge.clear(); pk.clear();
    maxdeg = 0; mindeg = 1 << 30; maxcdeg = 0;
    for (int j = 0; j < b; j++) {
        rp[j] = (int)col.size();
        for (int i = 0; i < c; i++) {
            int v = hd[(size_t)j * c + i];
            if (v == -1) continue;
            // rotate() reduces any shift mod M (decoders.cpp:335-339)
            int s = ((v % Z) + Z) % Z;
            int local = (int)col.size() - rp[j];
            col.push_back(i); sh.push_back(s); row.push_back(j);
            pk.push_back((uint32_t)i | ((uint32_t)s << 8) | ((uint32_t)local << 20));
        }
        int d = (int)col.size() - rp[j];
        if (d > maxdeg) maxdeg = d;
        if (d < mindeg) mindeg = d;
    }
    E = (int)col.size();
    rp[b] = E;
    for (int i = 0; i < c; i++) {
        cp[i] = (int)cedge.size();
        for (int e = 0; e < E; e++)
            if (col[e] == i) cedge.push_back(e);        // e ascends with the block row
        int d = (int)cedge.size() - cp[i];
        if (d > maxcdeg) maxcdeg = d;
    }
    cp[c] = E;
    all_cw_2 = 1;
    for (int i = 0; i < c; i++)
        if (cp[i + 1] - cp[i] != 2) { all_cw_2 = 0; break; }
    return E > 0;
}

} // namespace ldpcb200
