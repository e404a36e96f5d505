// SavedModel + tensor-bundle reader / writer (host only).  The two C-ABI entry points are declared in
// include/shifu_b200.h (sb_savedmodel_write / sb_savedmodel_read).
#pragma once
#include "../../include/shifu_b200.h"
