// K5: hard-negative mining (placeholder until the kernels land; see cc_mine.cu history).
#include "cc_internal.h"
namespace cc {
int mine_free(cc_ctx*) { return CC_OK; }
}
using namespace cc;
extern "C" {
int cc_bg_clear(cc_ctx*) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
int cc_bg_register(cc_ctx*, const uint8_t*, int, int) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
int cc_bg_count(cc_ctx*) { return 0; }
int cc_cascade_set_stages(cc_ctx*, const cc_stage*, int) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
int cc_mine_negatives(cc_ctx*, int, int, double, cc_reader_state*, int64_t*, int32_t*) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
int cc_mine_positives(cc_ctx*, const uint8_t*, int, int, int32_t*, int64_t*, int32_t*) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
int cc_mine_scan(cc_ctx*, int, int, int64_t*, int64_t*) { set_error("mining not built yet"); return CC_ERR_UNSUPPORTED; }
}
