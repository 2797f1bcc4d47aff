// Internal declarations shared between the .cu translation units (not part of the public C ABI).
#pragma once
#include "common.cuh"
#include "../../include/nemo_b200.h"
