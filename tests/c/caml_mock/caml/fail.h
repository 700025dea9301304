/* MOCK of <caml/fail.h> */
#ifndef CAML_MOCK_FAIL_H
#define CAML_MOCK_FAIL_H
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#endif
