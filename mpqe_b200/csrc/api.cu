// Library-level entry points: version, error text.
#include <stdarg.h>

#include "common.cuh"

namespace mpqe {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace mpqe

extern "C" const char* mpqe_b200_last_error(void) { return mpqe::g_error; }
extern "C" int mpqe_b200_version(void) { return 100; }
extern "C" int mpqe_b200_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(mpqe_term_t);
    case 1: return (int)sizeof(mpqe_layer_group_t);
    case 2: return (int)sizeof(mpqe_wgrad_dest_t);
    case 3: return (int)sizeof(mpqe_wgrad_operand_t);
    case 4: return (int)sizeof(mpqe_gather_item_t);
    case 5: return (int)sizeof(mpqe_margin_item_t);
    case 6: return (int)sizeof(mpqe_colsum_item_t);
    case 7: return (int)sizeof(mpqe_matsum_item_t);
    case 8: return (int)sizeof(mpqe_l2_item_t);
    case 9: return (int)sizeof(mpqe_adam_item_t);
    case 10: return (int)sizeof(mpqe_adam_table_t);
    default: return -1;
  }
}
