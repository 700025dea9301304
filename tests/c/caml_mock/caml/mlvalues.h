/* MOCK of <caml/mlvalues.h> -- declarations only, for a gcc -fsyntax-only pass over ocaml/mcmc_gpu_stubs.c in an
 * image without an OCaml toolchain (tests/test_abi.py::test_ocaml_stubs_parse).  NOT the OCaml runtime: types and
 * macros have the runtime's shapes so that type errors in the stubs are caught; nothing here can be linked. */
#ifndef CAML_MOCK_MLVALUES_H
#define CAML_MOCK_MLVALUES_H
#include <stdint.h>
typedef intptr_t value;
typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef uintnat mlsize_t;
#define Val_unit ((value)1)
#define Val_long(x) ((value)(((intnat)(x) << 1) + 1))
#define Long_val(x) ((intnat)(x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int)Long_val(x))
#define Is_block(x) (((x) & 1) == 0)
#define Field(x, i) (((value *)(x))[i])
#define Wosize_val(x) ((mlsize_t)(((uintnat *)(x))[-1] >> 10))
double caml_mock_double_val(value v);
#define Double_val(v) caml_mock_double_val(v)
int64_t caml_mock_int64_val(value v);
#define Int64_val(v) caml_mock_int64_val(v)
#define CAMLprim
#endif
