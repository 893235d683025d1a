/*
 * Element type of the dense blocks exchanged through spasm_schur_dense*() and
 * spasm_ffpack_rref(): the reference picks float / double / int64 from the
 * size of the prime because FFLAS-FFPACK computes in those types
 * (reference: src/spasm_ffpack.cpp:100-149).  The GPU path computes in int32
 * limbs whatever the prime; these helpers only convert at the API boundary so
 * that callers written against the reference keep working.
 */
#include <assert.h>
#include "spasm.h"

/* reference: src/spasm_ffpack.cpp:130-139 */
spasm_datatype spasm_datatype_choose(i64 prime)
{
	if (prime <= 8191)
		return SPASM_FLOAT;
	if (prime <= 189812531)
		return SPASM_DOUBLE;
	return SPASM_I64;
}

/* reference: src/spasm_ffpack.cpp:120-128 */
size_t spasm_datatype_size(spasm_datatype datatype)
{
	switch (datatype) {
	case SPASM_FLOAT:  return sizeof(float);
	case SPASM_DOUBLE: return sizeof(double);
	case SPASM_I64:    return sizeof(i64);
	}
	assert(false);
	return 0;
}

/* reference: src/spasm_ffpack.cpp:141-149 */
const char *spasm_datatype_name(spasm_datatype datatype)
{
	switch (datatype) {
	case SPASM_FLOAT:  return "float";
	case SPASM_DOUBLE: return "double";
	case SPASM_I64:    return "i64";
	}
	assert(false);
	return NULL;
}

/* reference: src/spasm_ffpack.cpp:100-108 */
spasm_ZZp spasm_datatype_read(const void *A, size_t i, spasm_datatype datatype)
{
	switch (datatype) {
	case SPASM_FLOAT:  return (spasm_ZZp) ((const float *) A)[i];
	case SPASM_DOUBLE: return (spasm_ZZp) ((const double *) A)[i];
	case SPASM_I64:    return (spasm_ZZp) ((const i64 *) A)[i];
	}
	assert(false);
	return 0;
}

/* reference: src/spasm_ffpack.cpp:110-118 */
void spasm_datatype_write(void *A, size_t i, spasm_datatype datatype, spasm_ZZp value)
{
	switch (datatype) {
	case SPASM_FLOAT:  ((float *) A)[i] = (float) value; return;
	case SPASM_DOUBLE: ((double *) A)[i] = (double) value; return;
	case SPASM_I64:    ((i64 *) A)[i] = value; return;
	}
	assert(false);
}
