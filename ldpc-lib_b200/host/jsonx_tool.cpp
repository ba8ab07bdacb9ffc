// `jsonx_rt <in.jsonx> <select-path or ""> <out.jsonx>`: parse, select, print -- the .jsonx boundary of the
// drop-in driver as a command-line tool (used by tests/test_settings.py against the reference's output).
#include <cstring>
#include "settings.h"

int main(int argc, char* argv[])
{
    if (argc != 4) { fprintf(stderr, "usage: %s <in.jsonx> <select-path or \"\"> <out.jsonx>\n", argv[0]); return 2; }
    settings s = settings::from_file(argv[1]);
    if (strlen(argv[2])) {
        settings t = s.select(argv[2]);
        t.to_file(argv[3]);
    } else
        s.to_file(argv[3]);
    return 0;
}
