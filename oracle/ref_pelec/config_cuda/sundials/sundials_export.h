/* Hand-written export header for a static, in-tree build of SUNDIALS 7.1.1 (oracle only). */
#ifndef SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT
#define SUNDIALS_NO_EXPORT
#define SUNDIALS_DEPRECATED __attribute__((__deprecated__))
#define SUNDIALS_DEPRECATED_EXPORT SUNDIALS_EXPORT SUNDIALS_DEPRECATED
#define SUNDIALS_DEPRECATED_NO_EXPORT SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED
#endif
