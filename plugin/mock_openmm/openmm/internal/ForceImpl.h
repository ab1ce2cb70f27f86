/* STAND-IN for openmm/internal/ForceImpl.h. */
#ifndef MOCK_OPENMM_FORCEIMPL_H_
#define MOCK_OPENMM_FORCEIMPL_H_
namespace OpenMM { class ForceImpl { public: virtual ~ForceImpl() {} }; }
#endif
