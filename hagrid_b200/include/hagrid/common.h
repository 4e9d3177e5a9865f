// Compatibility name for src/common.h: scalar helpers are in hgb_types.h; the
// only extra the front end needs from here is the device timer.
#ifndef COMMON_H
#define COMMON_H
#include <functional>
#include "hgb_types.h"

namespace hagrid {

/// Milliseconds the device spent on everything `work` enqueued on the legacy
/// default stream (cudaEvent pair; src/common.h:15, src/profile.cu:5-18).
HOST float profile(std::function<void()> work);

} // namespace hagrid
#endif
