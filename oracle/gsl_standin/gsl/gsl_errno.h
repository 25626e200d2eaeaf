/* TEST INFRASTRUCTURE ONLY - minimal stand-in for <gsl/gsl_errno.h>.
 * GSL (GNU Scientific Library) is the reference's un-vendored dependency
 * (README.md:37 "GSL 2.4"); it is absent from this image.  This header declares
 * only what the reference sources under /root/reference use. */
#ifndef AIRICE_GSL_STANDIN_ERRNO_H
#define AIRICE_GSL_STANDIN_ERRNO_H
#ifdef __cplusplus
extern "C" {
#endif
enum {
  GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2,
  GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5,
  GSL_EFACTOR = 6, GSL_ESANITY = 7, GSL_ENOMEM = 8, GSL_EBADFUNC = 9,
  GSL_ERUNAWAY = 10, GSL_EMAXITER = 11, GSL_EZERODIV = 12, GSL_EBADTOL = 13,
  GSL_ETOL = 14, GSL_EUNDRFLW = 15, GSL_EOVRFLW = 16, GSL_ELOSS = 17, GSL_EROUND = 18, GSL_EBADLEN = 19,
  GSL_ENOTSQR = 20, GSL_ESING = 21, GSL_EDIVERGE = 22
};
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
const char *gsl_strerror(const int gsl_errno);
#ifdef __cplusplus
}
#endif
#endif
