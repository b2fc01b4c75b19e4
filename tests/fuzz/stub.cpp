#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <stdint.h>
namespace ccsm {
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) { (void)fmt; }
}
