#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include "symbolic.h"
struct cb200_settings;
namespace cb200 {
void set_error(const std::string& s);
SymbolicOptions options_from_settings(const cb200_settings* st);
extern thread_local std::vector<int32_t> g_block_hint;
}
