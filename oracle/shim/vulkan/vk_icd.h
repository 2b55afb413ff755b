#pragma once
#include "vulkan_core.h"
// loader/ICD interface constant (public LunarG loader header value)
#define ICD_LOADER_MAGIC 0x01CDC0DE
