#pragma once
#include "la.cuh"
