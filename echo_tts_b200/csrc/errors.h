#pragma once
#include <cstdarg>
namespace echo {
void set_error(const char* fmt, ...);
}
