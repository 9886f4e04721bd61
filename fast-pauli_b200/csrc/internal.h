/* Private glue between the translation units of libfastpauli_b200.so (not part of the public C ABI). */
#ifndef FASTPAULI_B200_INTERNAL_H
#define FASTPAULI_B200_INTERNAL_H
#ifdef __cplusplus
extern "C"
{
#endif
    /* records `msg` as this thread's fp_last_error() and returns `code` */
    int fp_internal_set_error(int code, const char *msg);
#ifdef __cplusplus
}
#endif
#endif
