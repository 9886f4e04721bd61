// _fast_pauli -- native Python module over the C++ classes of fast_pauli.hpp.
//
// The reference ships a nanobind module of the same name (fast_pauli/cpp/src/fast_pauli.cpp:28-106 and
// src/include/__*_bindings.hpp); nanobind is not available here, pybind11 is.  The Python surface is the
// reference's: same classes, method names, keyword names, overloads, 1-D / 2-D dispatch, complex128-only arrays,
// zero-initialised outputs, ValueError for std::invalid_argument, pickling, `helpers` submodule.  All compute goes
// through the C++ classes, i.e. through libfastpauli_b200.so on the GPU; there is no CPU path.
//
// (fast-pauli_b200/__init__.py is the ctypes front-end over the same C ABI; it additionally accepts device-resident
// arrays and complex64.  Both are tested against the same cases, tests/test_reference_python_cases.py.)
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "fast_pauli.hpp"

namespace py = pybind11;
namespace fp = fast_pauli;
using namespace pybind11::literals;

using cd = std::complex<double>;
using CArray = py::array_t<cd>;
template <size_t N> using Span = std::mdspan<cd, std::dextents<size_t, N>>;

namespace
{

// ---- ndarray plumbing (reference: __nb_helpers.hpp:55-211) -----------------------------------------------------
// Row-major contiguity is REQUIRED (a strided view raises ValueError like the reference, NB:65-73); any other dtype
// is converted to complex128 the way nanobind's implicit conversion does for the reference (PY_PS:185,191).
CArray as_complex(py::handle obj, char const *what)
{
    py::array arr = py::array::ensure(obj);
    if (!arr)
        throw py::type_error(std::string(what) + ": expected an array");
    if (!(arr.flags() & py::array::c_style))
        throw std::invalid_argument(std::string(what) + ": ndarray MUST have C-style (row-major, contiguous) strides");
    CArray out = CArray::ensure(arr); // converts the dtype when needed
    if (!out)
        throw py::type_error(std::string(what) + ": cannot convert to complex128");
    return out;
}

py::array_t<double> as_real(py::handle obj, char const *what)
{
    py::array arr = py::array::ensure(obj);
    if (!arr)
        throw py::type_error(std::string(what) + ": expected an array");
    if (!(arr.flags() & py::array::c_style))
        throw std::invalid_argument(std::string(what) + ": ndarray MUST have C-style (row-major, contiguous) strides");
    auto out = py::array_t<double>::ensure(arr);
    if (!out)
        throw py::type_error(std::string(what) + ": cannot convert to float64");
    return out;
}

CArray zeros(std::vector<py::ssize_t> const &shape)
{
    CArray a(shape);
    std::fill_n(a.mutable_data(), static_cast<size_t>(a.size()), cd(0));
    return a;
}

void need_1d_or_2d(CArray const &a, char const *what)
{
    if (a.ndim() != 1 && a.ndim() != 2)
        throw std::invalid_argument(std::string(what) + ": expected 1 or 2 dimensions, got " + std::to_string(a.ndim()));
}

// The C++ methods take non-const mdspans for their inputs as well (like the reference's); inputs are never written,
// so read-only numpy arrays (e.g. views of constants, np.broadcast_to results made contiguous) are accepted.
cd *ptr(CArray &a)
{
    return const_cast<cd *>(a.data());
}
Span<1> span1(CArray &a)
{
    return Span<1>(ptr(a), static_cast<size_t>(a.shape(0)));
}
Span<2> span2(CArray &a)
{
    return Span<2>(ptr(a), static_cast<size_t>(a.shape(0)), static_cast<size_t>(a.shape(1)));
}
// a 1-D state viewed as a (dim, 1) batch
Span<2> column(CArray &a)
{
    return Span<2>(ptr(a), static_cast<size_t>(a.shape(0)), 1);
}

std::vector<std::string> as_strings(std::vector<fp::PauliString> const &ps)
{
    std::vector<std::string> out(ps.size());
    std::transform(ps.begin(), ps.end(), out.begin(), [](fp::PauliString const &p) { return p.str(); });
    return out;
}

using Op = fp::PauliOp<double>;
using Sop = fp::SummedPauliOp<double>;

Op plus(Op const &lhs, Op const &rhs, cd sign)
{
    Op out(lhs);
    Op scaled(rhs);
    scaled.scale(sign);
    out.extend(scaled);
    return out;
}
Op plus(Op const &lhs, fp::PauliString const &rhs, cd sign)
{
    Op out(lhs);
    out.extend(rhs, sign, true);
    return out;
}

Sop make_sop(std::vector<fp::PauliString> const &strings, py::handle coeffs)
{
    CArray c = as_complex(coeffs, "coeffs");
    if (c.ndim() != 2)
        throw std::invalid_argument("coeffs must be a 2-D array of shape (n_pauli_strings, n_operators)");
    return Sop(strings, span2(c));
}

} // namespace

PYBIND11_MODULE(_fast_pauli, m)
{
    m.doc() = "B200-native fast-pauli hot path: native bindings over the C++ classes (GPU only, no CPU fallback)";

    // ------------------------------------------------------------------------------------------------ Pauli
    py::class_<fp::Pauli>(m, "Pauli")
        .def(py::init<>())
        .def(py::init([](int code) { return fp::Pauli(code); }), "code"_a)
        .def(py::init([](py::str const &symbol) {
                 std::string s = symbol;
                 if (s.size() != 1) // the reference's overload takes a `char`: a longer str matches no overload
                     throw py::type_error("Pauli(symbol): expected a single character");
                 return fp::Pauli(s[0]);
             }),
             "symbol"_a)
        .def("__matmul__", [](fp::Pauli const &a, fp::Pauli const &b) { return a * b; }, py::is_operator())
        .def("to_tensor",
             [](fp::Pauli const &self) {
                 CArray out = zeros({2, 2});
                 self.to_tensor(span2(out));
                 return out;
             })
        .def("clone", [](fp::Pauli const &self) { return fp::Pauli(self); })
        .def("__str__", [](fp::Pauli const &self) { return std::string(1, self.symbol()); })
        .def("__repr__", [](fp::Pauli const &self) { return std::string("Pauli('") + self.symbol() + "')"; })
        .def("__eq__", [](fp::Pauli const &a, fp::Pauli const &b) { return a.code == b.code; }, py::is_operator())
        .def("__hash__", [](fp::Pauli const &self) { return static_cast<size_t>(self.code); })
        .def(py::pickle([](fp::Pauli const &self) { return static_cast<int>(self.code); },
                        [](int code) { return fp::Pauli(code); }));

    // ------------------------------------------------------------------------------------------------ PauliString
    py::class_<fp::PauliString>(m, "PauliString")
        .def(py::init<>())
        .def(py::init<std::string const &>(), "string"_a)
        .def(py::init([](std::vector<fp::Pauli> paulis) { return fp::PauliString(std::move(paulis)); }), "paulis"_a)
        .def("__str__", &fp::PauliString::str)
        .def("__repr__", [](fp::PauliString const &self) { return "PauliString(\"" + self.str() + "\")"; })
        .def("__eq__", [](fp::PauliString const &a, fp::PauliString const &b) { return a.paulis == b.paulis; },
             py::is_operator())
        .def("__hash__", [](fp::PauliString const &self) { return std::hash<fp::PauliString>()(self); })
        .def("__matmul__", [](fp::PauliString const &a, fp::PauliString const &b) { return a * b; }, py::is_operator())
        .def("__add__",
             [](fp::PauliString const &a, fp::PauliString const &b) {
                 Op out(std::vector<cd>{1.0}, {a});
                 out.extend(b, 1.0, true);
                 return out;
             },
             py::is_operator())
        .def("__sub__",
             [](fp::PauliString const &a, fp::PauliString const &b) {
                 Op out(std::vector<cd>{1.0}, {a});
                 out.extend(b, -1.0, true);
                 return out;
             },
             py::is_operator())
        .def_property_readonly("n_qubits", &fp::PauliString::n_qubits)
        .def_property_readonly("dim", &fp::PauliString::dim)
        .def_property_readonly("weight", [](fp::PauliString const &self) { return static_cast<int>(self.weight); })
        .def(
            "apply",
            [](fp::PauliString const &self, py::handle states_in, cd coeff) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "apply");
                std::vector<py::ssize_t> shape(states.shape(), states.shape() + states.ndim());
                CArray out = zeros(shape);
                if (states.ndim() == 1)
                    self.apply(std::execution::par, span1(out), span1(states)); // no coeff: B_PS:142 drops it too
                else
                    self.apply_batch(std::execution::par, span2(out), span2(states), coeff);
                return out;
            },
            "states"_a, "coeff"_a = cd(1.0))
        .def(
            "expectation_value",
            [](fp::PauliString const &self, py::handle states_in, cd coeff) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "expectation_value");
                bool const one = states.ndim() == 1;
                CArray out = zeros({one ? py::ssize_t(1) : states.shape(1)});
                self.expectation_value(std::execution::par, span1(out), one ? column(states) : span2(states), coeff);
                return out;
            },
            "states"_a, "coeff"_a = cd(1.0))
        .def("to_tensor",
             [](fp::PauliString const &self) {
                 auto const d = static_cast<py::ssize_t>(self.dim());
                 CArray out = zeros({d, d});
                 if (d)
                     self.to_tensor(span2(out));
                 return out;
             })
        .def("clone", [](fp::PauliString const &self) { return fp::PauliString(self); })
        .def(py::pickle([](fp::PauliString const &self) { return self.str(); },
                        [](std::string const &s) { return fp::PauliString(s); }));

    // ------------------------------------------------------------------------------------------------ PauliOp
    py::class_<Op>(m, "PauliOp")
        .def(py::init<>())
        .def(py::init<std::vector<std::string> const &>(), "pauli_strings"_a)
        .def(py::init<std::vector<fp::PauliString>>(), "pauli_strings"_a)
        .def(py::init([](std::vector<cd> coefficients, std::vector<fp::PauliString> strings) {
                 return Op(std::move(coefficients), std::move(strings));
             }),
             "coefficients"_a, "pauli_strings"_a)
        .def(py::init([](std::vector<cd> coefficients, std::vector<std::string> const &strings) {
                 std::vector<fp::PauliString> ps(strings.begin(), strings.end());
                 return Op(std::move(coefficients), std::move(ps));
             }),
             "coefficients"_a, "pauli_strings"_a)
        .def("__matmul__", [](Op const &a, Op const &b) { return a * b; }, py::is_operator())
        .def("__matmul__", [](Op const &a, fp::PauliString const &b) { return a * b; }, py::is_operator())
        .def("__rmatmul__", [](Op const &self, fp::PauliString const &lhs) { return lhs * self; }, py::is_operator())
        .def("__mul__",
             [](Op const &self, cd factor) {
                 Op out(self);
                 out.scale(factor);
                 return out;
             },
             py::is_operator())
        .def("__rmul__",
             [](Op const &self, cd factor) {
                 Op out(self);
                 out.scale(factor);
                 return out;
             },
             py::is_operator())
        .def("__imul__",
             [](Op &self, cd factor) -> Op & {
                 self.scale(factor);
                 return self;
             },
             py::is_operator())
        .def("__add__", [](Op const &a, Op const &b) { return plus(a, b, 1.0); }, py::is_operator())
        .def("__add__", [](Op const &a, fp::PauliString const &b) { return plus(a, b, 1.0); }, py::is_operator())
        .def("__radd__", [](Op const &self, fp::PauliString const &lhs) { return plus(self, lhs, 1.0); },
             py::is_operator())
        .def("__iadd__",
             [](Op &self, Op const &other) -> Op & {
                 self.extend(other);
                 return self;
             },
             py::is_operator())
        .def("__iadd__",
             [](Op &self, fp::PauliString const &other) -> Op & {
                 self.extend(other, 1.0, true);
                 return self;
             },
             py::is_operator())
        .def("__sub__", [](Op const &a, Op const &b) { return plus(a, b, -1.0); }, py::is_operator())
        .def("__sub__", [](Op const &a, fp::PauliString const &b) { return plus(a, b, -1.0); }, py::is_operator())
        .def("__rsub__",
             [](Op const &self, fp::PauliString const &lhs) {
                 Op out(self);
                 out.scale(cd(-1.0));
                 out.extend(lhs, 1.0, true);
                 return out;
             },
             py::is_operator())
        .def("__isub__",
             [](Op &self, Op const &other) -> Op & {
                 Op neg(other);
                 neg.scale(cd(-1.0));
                 self.extend(neg);
                 return self;
             },
             py::is_operator())
        .def("__isub__",
             [](Op &self, fp::PauliString const &other) -> Op & {
                 self.extend(other, -1.0, true);
                 return self;
             },
             py::is_operator())
        .def("extend", [](Op &self, Op const &other) { self.extend(other); }, "other"_a)
        .def(
            "extend",
            [](Op &self, fp::PauliString const &other, cd multiplier, bool dedupe) {
                self.extend(other, multiplier, dedupe);
            },
            "other"_a, "multiplier"_a, "dedupe"_a = true)
        .def_property_readonly("dim", &Op::dim)
        .def_property_readonly("n_qubits", &Op::n_qubits)
        .def_property_readonly("n_pauli_strings", &Op::n_pauli_strings)
        .def_property_readonly("coeffs", [](Op const &self) { return self.coeffs; })
        .def_property_readonly("pauli_strings", [](Op const &self) { return self.pauli_strings; })
        .def_property_readonly("pauli_strings_as_str", [](Op const &self) { return as_strings(self.pauli_strings); })
        .def("scale", [](Op &self, cd factor) { self.scale(factor); }, "factor"_a)
        .def(
            "scale",
            [](Op &self, py::handle factors_in) {
                CArray f = as_complex(factors_in, "factors");
                if (f.ndim() != 1)
                    throw std::invalid_argument("factors must be a 1-D array");
                self.scale(span1(f));
            },
            "factors"_a)
        .def(
            "apply",
            [](Op const &self, py::handle states_in) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "apply");
                std::vector<py::ssize_t> shape(states.shape(), states.shape() + states.ndim());
                CArray out = zeros(shape);
                if (states.ndim() == 1)
                    self.apply(std::execution::par, span1(out), span1(states));
                else
                    self.apply(std::execution::par, span2(out), span2(states));
                return out;
            },
            "states"_a)
        .def(
            "expectation_value",
            [](Op const &self, py::handle states_in) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "expectation_value");
                bool const one = states.ndim() == 1;
                CArray out = zeros({one ? py::ssize_t(1) : states.shape(1)});
                self.expectation_value(std::execution::par, span1(out), one ? column(states) : span2(states));
                return out;
            },
            "states"_a)
        .def("to_tensor",
             [](Op const &self) {
                 auto const d = static_cast<py::ssize_t>(self.dim());
                 CArray out = zeros({d, d});
                 if (d)
                     self.to_tensor(span2(out));
                 return out;
             })
        .def("clone", [](Op const &self) { return Op(self); })
        .def(py::pickle([](Op const &self) { return py::make_tuple(self.coeffs, as_strings(self.pauli_strings)); },
                        [](py::tuple t) {
                            auto strings = t[1].cast<std::vector<std::string>>();
                            std::vector<fp::PauliString> ps(strings.begin(), strings.end());
                            return Op(t[0].cast<std::vector<cd>>(), std::move(ps));
                        }));

    // ------------------------------------------------------------------------------------------------ SummedPauliOp
    py::class_<Sop>(m, "SummedPauliOp")
        .def(py::init<>())
        .def(py::init([](std::vector<std::string> const &strings, py::handle coeffs) {
                 std::vector<fp::PauliString> ps(strings.begin(), strings.end());
                 return make_sop(ps, coeffs);
             }),
             "pauli_strings"_a, "coeffs"_a)
        .def(py::init([](std::vector<fp::PauliString> const &strings, py::handle coeffs) {
                 return make_sop(strings, coeffs);
             }),
             "pauli_strings"_a, "coeffs"_a)
        .def_property_readonly("dim", &Sop::dim)
        .def_property_readonly("n_qubits", &Sop::n_qubits)
        .def_property_readonly("n_operators", &Sop::n_operators)
        .def_property_readonly("n_pauli_strings", &Sop::n_pauli_strings)
        .def_property(
            "coeffs",
            // (n_operators, n_pauli_strings): the reference's getter / setter use the transpose of the constructor's
            // orientation (B_SPO:113-135)
            [](Sop const &self) {
                auto const S = static_cast<py::ssize_t>(self.n_pauli_strings());
                auto const K = static_cast<py::ssize_t>(self.n_operators());
                CArray out = zeros({K, S});
                auto o = out.mutable_unchecked<2>();
                for (py::ssize_t i = 0; i < S; ++i)
                    for (py::ssize_t k = 0; k < K; ++k)
                        o(k, i) = self.coeffs(static_cast<size_t>(i), static_cast<size_t>(k));
                return out;
            },
            [](Sop &self, py::handle coeffs_new) {
                CArray c = as_complex(coeffs_new, "coeffs");
                if (c.ndim() != 2 || static_cast<size_t>(c.shape(0)) != self.n_operators() ||
                    static_cast<size_t>(c.shape(1)) != self.n_pauli_strings())
                    throw std::invalid_argument(
                        "The shape of provided coeffs must match the number of operators and PauliStrings");
                std::vector<cd> flat(self.n_pauli_strings() * self.n_operators());
                auto r = c.unchecked<2>();
                for (size_t i = 0; i < self.n_pauli_strings(); ++i)
                    for (size_t k = 0; k < self.n_operators(); ++k)
                        flat[i * self.n_operators() + k] = r(static_cast<py::ssize_t>(k), static_cast<py::ssize_t>(i));
                self = Sop(self.pauli_strings, flat); // rebuilds the device plan on next use
            })
        .def_property_readonly("pauli_strings", [](Sop const &self) { return self.pauli_strings; })
        .def_property_readonly("pauli_strings_as_str", [](Sop const &self) { return as_strings(self.pauli_strings); })
        .def(
            "apply",
            [](Sop const &self, py::handle states_in) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "apply");
                std::vector<py::ssize_t> shape(states.shape(), states.shape() + states.ndim());
                CArray out = zeros(shape);
                if (states.ndim() == 1)
                    self.apply(std::execution::par, column(out), column(states));
                else
                    self.apply(std::execution::par, span2(out), span2(states));
                return out;
            },
            "states"_a)
        .def(
            "apply_weighted",
            [](Sop const &self, py::handle states_in, py::handle data_in) {
                CArray states = as_complex(states_in, "states");
                py::array_t<double> data = as_real(data_in, "data");
                need_1d_or_2d(states, "apply_weighted");
                if (data.ndim() != states.ndim())
                    throw std::invalid_argument("states and data must have the same number of dimensions");
                std::vector<py::ssize_t> shape(states.shape(), states.shape() + states.ndim());
                CArray out = zeros(shape);
                bool const one = states.ndim() == 1;
                std::mdspan<double, std::dextents<size_t, 2>> w(const_cast<double *>(data.data()),
                                                                static_cast<size_t>(data.shape(0)),
                                                                one ? 1 : static_cast<size_t>(data.shape(1)));
                self.apply_weighted(std::execution::par, one ? column(out) : span2(out),
                                    one ? column(states) : span2(states), w);
                return out;
            },
            "states"_a, "data"_a)
        .def(
            "expectation_value",
            [](Sop const &self, py::handle states_in) {
                CArray states = as_complex(states_in, "states");
                need_1d_or_2d(states, "expectation_value");
                bool const one = states.ndim() == 1;
                auto const K = static_cast<py::ssize_t>(self.n_operators());
                CArray out = one ? zeros({K}) : zeros({K, states.shape(1)});
                Span<2> o(ptr(out), static_cast<size_t>(K), one ? 1 : static_cast<size_t>(states.shape(1)));
                self.expectation_value(std::execution::par, o, one ? column(states) : span2(states));
                return out;
            },
            "states"_a)
        .def("to_tensor",
             [](Sop const &self) {
                 auto const d = static_cast<py::ssize_t>(self.dim());
                 CArray out = zeros({static_cast<py::ssize_t>(self.n_operators()), d, d});
                 self.to_tensor(Span<3>(ptr(out), self.n_operators(), self.dim(), self.dim()));
                 return out;
             })
        .def("clone", [](Sop const &self) { return Sop(self); })
        .def("split", &Sop::split)
        .def("square", &Sop::square)
        .def(py::pickle(
            [](Sop const &self) {
                return py::make_tuple(as_strings(self.pauli_strings), self.coeffs_raw, self.n_operators());
            },
            [](py::tuple t) {
                auto strings = t[0].cast<std::vector<std::string>>();
                std::vector<fp::PauliString> ps(strings.begin(), strings.end());
                return Sop(ps, t[1].cast<std::vector<cd>>());
            }));

    // ------------------------------------------------------------------------------------------------ helpers
    auto helpers = m.def_submodule("helpers");
    helpers.def("get_nontrivial_paulis", &fp::get_nontrivial_paulis, "weight"_a);
    helpers.def("calculate_pauli_strings", &fp::calculate_pauli_strings, "n_qubits"_a, "weight"_a);
    helpers.def("calculate_pauli_strings_max_weight", &fp::calculate_pauli_strings_max_weight, "n_qubits"_a,
                "weight"_a);
    helpers.def("pauli_string_sparse_repr", &fp::get_sparse_repr<double>, "paulis"_a);
}
