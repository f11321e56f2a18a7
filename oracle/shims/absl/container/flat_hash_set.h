#pragma once
#include "flat_hash_map.h"
