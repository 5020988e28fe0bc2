// lmcma_kernels.cuh — the kernels of one LM-CMA generation (see lmcma_common.cuh).
#pragma once
#include "k_cost.cuh"
#include "k_rank.cuh"
#include "k_sample.cuh"
#include "k_update.cuh"
#include "k_prior.cuh"
#include "k_edt.cuh"
#include "k_gram.cuh"
