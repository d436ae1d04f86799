/* TEST INFRASTRUCTURE: host build of bithtm_b200/csrc/np_expf.h so the sequence
 * the CUDA boost kernel uses can be checked against np.exp on the CPU
 * (regularizations.py:16).  Built by oracle/Makefile into oracle/_build/. */
#include <stddef.h>
#include "../bithtm_b200/csrc/np_expf.h"

void bh_np_expf_array(const float* x, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = bh_np_expf(x[i]);
}
