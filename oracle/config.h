/* TEST INFRASTRUCTURE ONLY: hand-written stand-in for the autoconf-generated config.h of the
 * reference build (configure.ac), used only when compiling the reference sources into oracle/_ref. */
#define HAVE_CXX11 1
#define USE_OPENMP 1
#define HAVE_ISNAN 1
#define HAVE_ISINF 1
#define HAVE_ERF 1
#define HAVE_ERFC 1
#define HAVE_COPYSIGN 1
#define HAVE_EXECINFO_H 1
#define HAVE_BACKTRACE 1
#define F77_FUNC(name,NAME) name ## _
#define F77_FUNC_(name,NAME) name ## _
#define PACKAGE_VERSION "0.96"
#define PACKAGE_STRING "scuff-em 0.96"
#define VERSION "0.96"
