/* MOCK of <caml/memory.h> (see mlvalues.h in this directory) */
#ifndef CAML_MOCK_MEMORY_H
#define CAML_MOCK_MEMORY_H
#include "mlvalues.h"
#define CAMLparam0() int caml__frame = 0; (void)caml__frame
#define CAMLparam1(a) CAMLparam0(); (void)(a)
#define CAMLparam2(a, b) CAMLparam1(a); (void)(b)
#define CAMLparam3(a, b, c) CAMLparam2(a, b); (void)(c)
#define CAMLparam4(a, b, c, d) CAMLparam3(a, b, c); (void)(d)
#define CAMLparam5(a, b, c, d, e) CAMLparam4(a, b, c, d); (void)(e)
#define CAMLxparam1(a) (void)(a)
#define CAMLxparam2(a, b) (void)(a); (void)(b)
#define CAMLxparam3(a, b, c) CAMLxparam2(a, b); (void)(c)
#define CAMLxparam4(a, b, c, d) CAMLxparam3(a, b, c); (void)(d)
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLlocal3(a, b, c) value a = Val_unit, b = Val_unit, c = Val_unit
#define CAMLreturn(x) return (x)
void caml_mock_store_field(value block, int i, value v);
#define Store_field(b, i, v) caml_mock_store_field(b, i, v)
#endif
