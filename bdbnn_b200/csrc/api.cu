// Library-level entry points: version, thread-local error text.
#include <stdarg.h>

#include "common.cuh"

namespace bdbnn {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
}

}  // namespace bdbnn

extern "C" int bdbnn_version(void) { return 1000; }

extern "C" const char* bdbnn_last_error_string(void) { return bdbnn::last_error_buf(); }
