// TEMPORARY: construction stages not yet implemented in this checkpoint.
#include <cstdio>
#include <cstdlib>
#include "build.h"
namespace hagrid {
static void missing(const char* what) { std::fprintf(stderr, "hagrid_b200: %s not implemented yet\n", what); std::abort(); }
void build_grid(MemManager&, const Tri*, int, Grid&, float, float) { missing("build_grid"); }
void merge_grid(MemManager&, Grid&, float) { missing("merge_grid"); }
void flatten_grid(MemManager&, Grid&) { missing("flatten_grid"); }
void expand_grid(MemManager&, Grid&, const Tri*, int) { missing("expand_grid"); }
bool compress_grid(MemManager&, Grid&) { missing("compress_grid"); return false; }
}
