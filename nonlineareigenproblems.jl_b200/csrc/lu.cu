// placeholder until the device LU lands (next milestone)
#include "common.h"
namespace nepb {
void lu_symbolic_release(void*) {}
}  // namespace nepb
