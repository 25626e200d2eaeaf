/* TEST INFRASTRUCTURE ONLY - stand-in for <gsl/gsl_fit.h>; the reference includes
 * it (MultiRayAirIceRefraction.h:18) but calls nothing from it. */
#ifndef AIRICE_GSL_STANDIN_FIT_H
#define AIRICE_GSL_STANDIN_FIT_H
#endif
