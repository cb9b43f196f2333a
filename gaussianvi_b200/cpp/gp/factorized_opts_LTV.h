// gp/factorized_opts_LTV.h:9-14 of the reference: LTVGpPrior over LTV_GP
#pragma once
#include "../gvi/gvi.h"
