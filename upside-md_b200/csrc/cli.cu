// upside_main: placeholder until the batched MD driver lands (see main_cli in a later commit)
#include <cstdio>
#include "../../include/engine_c_library.h"
extern "C" int upside_main(int argc, const char* const* argv, int verbose) {
    (void)argc; (void)argv; (void)verbose;
    fprintf(stderr, "ERROR: upside_main is not available in this build\n");
    return 1;
}
