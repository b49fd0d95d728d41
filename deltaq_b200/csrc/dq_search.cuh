#pragma once
#include "dq_common.cuh"
