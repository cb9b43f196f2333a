// forwarding header: the reference include path resolves to the gvib200 facade
#pragma once
#include "../gvi/gvi.h"
