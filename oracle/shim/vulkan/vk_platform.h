#pragma once
#include <stdint.h>
#include <stddef.h>
