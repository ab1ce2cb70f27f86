/* STAND-IN for openmm/TabulatedFunction.h (compile check only). */
#ifndef MOCK_OPENMM_TABULATEDFUNCTION_H_
#define MOCK_OPENMM_TABULATEDFUNCTION_H_
namespace OpenMM { class TabulatedFunction { public: virtual ~TabulatedFunction() {} }; }
#endif
