#pragma once
#include <vector>
#include "spasm_b200.h"

namespace sb {

struct Stats {
	struct spasm_b200_stats pub;
	std::vector<int> pair_row, pair_col, pair_start;
	Stats() { memset(&pub, 0, sizeof(pub)); }
};

/* count one kernel launch of this library (the "gpu_launches" claim of bench.py) */
#define LAUNCHED(n) (sb::stats().pub.kernel_launches += (n))

}  // namespace sb
