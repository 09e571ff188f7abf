/* fo_sa.c -- Spalart-Allmaras source and eddy viscosity (TEST INFRASTRUCTURE, see fasts_oracle.h). */
#include <math.h>
#include "fo_internal.h"
#include "fasts_oracle.h"
