/* shim_host.c — weak defaults for the two symbols synth.c's replacement imports
 * from its host program (SURVEY §8b "Symbols the path imports"): `int debug`
 * (skred.h:102, defined by skred.c) and `mw_free` (miniwav.h:51, miniwav.c).
 * Weak, so that inside a real skred build the program's own definitions win;
 * they only make libskred_shim_v<N>.so loadable on its own (bench, tests). */
#include <stdlib.h>

__attribute__((weak)) int debug = 0;

__attribute__((weak)) float *mw_free(float *f) {
  free(f);
  return NULL;
}
