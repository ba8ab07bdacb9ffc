// TEST INFRASTRUCTURE ONLY.  `jsonx_rt <in.jsonx> <select-path or ""> <out.jsonx>`: the UNMODIFIED
// reference's settings::from_file -> select -> to_file (settings.cpp:26-44, 364-399, 507-515), linked from
// the reference's own object files by oracle/Makefile; used by tests/golden/make_golden.py.
#include <cstring>
#include "settings.h"

int main(int argc, char* argv[])
{
    if (argc != 4) return 2;
    settings s = settings::from_file(argv[1]);
    if (strlen(argv[2])) {
        settings t = s.select(argv[2]);
        t.to_file(argv[3]);
    } else
        s.to_file(argv[3]);
    return 0;
}
