// TEST INFRASTRUCTURE ONLY -- never linked into libbn_b200.so.
//
// Builds the reference's legacy stand-alone program
// (/root/reference/Bayes-networks/main.cpp with random4f.h, cholesky21.h)
// UNMODIFIED: the file is #included where it lies; only its five absolute
// fopen() paths (main.cpp:53,64,344-346) are redirected by a macro so inputs
// are read from $BN_LEGACY_IN and outputs written to $BN_LEGACY_OUT.
// Output: oracle/_ref/legacy_main (git-ignored).  Used to reproduce the
// 1,100-row golden trace `iterations - null start.xlsx`.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

static FILE* bn_redirect_fopen(const char* path, const char* mode) {
  const char* base = strrchr(path, '/');
  base = base ? base + 1 : path;
  const char* dir = getenv(mode[0] == 'r' ? "BN_LEGACY_IN" : "BN_LEGACY_OUT");
  std::string p = std::string(dir ? dir : ".") + "/" + base;
  return fopen(p.c_str(), mode);
}
#define fopen bn_redirect_fopen
#include "main.cpp"
