#pragma once
#include "common.cuh"
#include "zp.cuh"

namespace sb {

struct PivotCounts { int fl, flcol, greedy; };

/* pinv (n) / qinv (m) are device arrays, overwritten.  reference: src/spasm_pivots.c:305-319 */
PivotCounts pivots_find(const DevCsr &A, int *d_pinv, int *d_qinv, bool greedy);

/* append rows d_rows[0:npiv] of A to U, pivot (column pinv[row]) first and scaled to 1; Uqinv updated.
 * reference: src/spasm_pivots.c:405-443 */
void append_pivotal_rows(const DevCsr &A, const int *d_rows, int npiv, const int *d_pinv, DevCsr &U, DevBuf<int> &Uqinv);

}  // namespace sb
