/* Bodies of the stand-in OpenMM (plugin/mock_openmm): the platform registry, Platform::createKernel, and a Context that owns
 * one ForceImpl per Force and sums calcForcesAndEnergy over the requested groups on the Reference platform's PlatformData.
 * SlicedNonbondedForce gets the reference's OWN SlicedNonbondedForceImpl (compiled unchanged); a plain NonbondedForce is
 * evaluated as a one-subset SlicedNonbondedForce through the same kernel (OpenMM's own nonbonded kernels are not here);
 * HarmonicBondForce is evaluated in place.  Test infrastructure of plugin/Makefile's run targets. */
#include "SlicedNonbondedForce.h"
#include "internal/SlicedNonbondedForceImpl.h"

#include "openmm/Context.h"
#include "openmm/Force.h"
#include "openmm/HarmonicBondForce.h"
#include "openmm/Kernel.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/internal/AssertionUtilities.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"

#include <cmath>
#include <vector>

namespace OpenMM {

/* ---- platforms ---- */
static std::vector<Platform*>& platforms() {
    static std::vector<Platform*> list;
    if (list.empty())
        list.push_back(new ReferencePlatform());        /* OpenMM always has its Reference platform */
    return list;
}

int Platform::getNumPlatforms() { return (int) platforms().size(); }
Platform& Platform::getPlatform(int index) { return *platforms().at(index); }
Platform& Platform::getPlatformByName(const std::string& name) {
    for (size_t i = 0; i < platforms().size(); i++)
        if (platforms()[i]->getName() == name)
            return *platforms()[i];
    throw OpenMMException("There is no registered Platform called \""+name+"\"");
}
void Platform::registerPlatform(Platform* platform) { platforms().push_back(platform); }
const std::string& Platform::getName() const {
    static const std::string name = "Platform";
    return name;
}
const std::string& Platform::getPropertyDefaultValue(const std::string&) const {
    static const std::string none;
    return none;
}
void Platform::registerKernelFactory(const std::string& name, KernelFactory* factory) {
    factories[name] = factory;      /* a later registration under the same kernel name replaces the earlier one */
}
KernelImpl* Platform::createKernelImpl(const std::string& name, ContextImpl& context) const {
    std::map<std::string, KernelFactory*>::const_iterator f = factories.find(name);
    if (f == factories.end())
        throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return f->second->createKernelImpl(name, *this, context);
}
Kernel Platform::createKernel(const std::string& name, ContextImpl& context) const {
    return Kernel(createKernelImpl(name, context));
}

void throwException(const char* file, int line, const std::string& details) {
    std::stringstream message;
    message << "Assertion failure at " << file << ":" << line << ".  " << details;
    throw OpenMMException(message.str());
}

/* ---- ForceImpls of the forces OpenMM itself would evaluate ---- */
class PlainNonbondedForceImpl : public ForceImpl {
public:
    PlainNonbondedForceImpl(const NonbondedForce& owner) : owner(owner), standIn(NULL), impl(NULL) {}
    ~PlainNonbondedForceImpl() { delete impl; delete standIn; }
    void initialize(ContextImpl& context) {
        delete impl; delete standIn;
        standIn = new OpenMMLab::SlicedNonbondedForce(owner, 1);
        impl = new OpenMMLab::SlicedNonbondedForceImpl(*standIn);
        impl->initialize(context);
    }
    const Force& getOwner() const { return owner; }
    double calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool includeEnergy, int groups) {
        return impl->calcForcesAndEnergy(context, includeForces, includeEnergy, groups);
    }
    std::map<std::string, double> getDefaultParameters() {
        std::map<std::string, double> parameters;
        for (int i = 0; i < owner.getNumGlobalParameters(); i++)
            parameters[owner.getGlobalParameterName(i)] = owner.getGlobalParameterDefaultValue(i);
        return parameters;
    }
    std::vector<std::string> getKernelNames() { return std::vector<std::string>(); }
    void updateParametersInContext(ContextImpl& context) {
        /* the stand-in is rebuilt from the owner's current parameters; the kernel keeps its handle (copyParametersToContext) */
        OpenMMLab::SlicedNonbondedForce* fresh = new OpenMMLab::SlicedNonbondedForce(owner, 1);
        for (int i = 0; i < fresh->getNumParticles(); i++) {
            double q, sigma, epsilon;
            fresh->getParticleParameters(i, q, sigma, epsilon);
            standIn->setParticleParameters(i, q, sigma, epsilon);
        }
        for (int i = 0; i < fresh->getNumExceptions(); i++) {
            int p1, p2;
            double qq, sigma, epsilon;
            fresh->getExceptionParameters(i, p1, p2, qq, sigma, epsilon);
            standIn->setExceptionParameters(i, p1, p2, qq, sigma, epsilon);
        }
        delete fresh;
        impl->updateParametersInContext(context);
    }
private:
    const NonbondedForce& owner;
    OpenMMLab::SlicedNonbondedForce* standIn;
    OpenMMLab::SlicedNonbondedForceImpl* impl;
};

ForceImpl* NonbondedForce::createImpl() const { return new PlainNonbondedForceImpl(*this); }

class HarmonicBondForceImpl : public ForceImpl {
public:
    HarmonicBondForceImpl(const HarmonicBondForce& owner) : owner(owner) {}
    void initialize(ContextImpl&) {}
    const Force& getOwner() const { return owner; }
    double calcForcesAndEnergy(ContextImpl& context, bool includeForces, bool, int groups) {
        if ((groups&(1<<owner.getForceGroup())) == 0)
            return 0.0;
        ReferencePlatform::PlatformData& data = *reinterpret_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
        double energy = 0;
        for (int i = 0; i < owner.getNumBonds(); i++) {
            const HarmonicBondForce::Bond& b = owner.getBond(i);
            const Vec3 delta = (*data.positions)[b.p2]-(*data.positions)[b.p1];
            const double r = std::sqrt(delta.dot(delta)), dr = r-b.length;
            energy += 0.5*b.k*dr*dr;
            if (includeForces && r > 0) {
                const Vec3 f = delta*(b.k*dr/r);
                (*data.forces)[b.p1] += f;
                (*data.forces)[b.p2] -= f;
            }
        }
        return energy;
    }
    std::vector<std::string> getKernelNames() { return std::vector<std::string>(); }
private:
    const HarmonicBondForce& owner;
};

ForceImpl* HarmonicBondForce::createImpl() const { return new HarmonicBondForceImpl(*this); }

/* ---- Context ---- */
struct Context::Data {
    const System* system;
    Platform* platform;
    std::vector<Vec3> positions, velocities, forces;
    Vec3 box[3];
    std::map<std::string, double> derivatives;
    ReferencePlatform::PlatformData platformData;
    ContextImpl* impl;
    std::vector<ForceImpl*> forceImpls;
};

Context::Context(const System& system, Integrator&, Platform& platform) : data(new Data) {
    data->system = &system;
    data->platform = &platform;
    data->impl = NULL;
    build();
}

Context::Context(const System& system, Integrator&, Platform& platform, const std::map<std::string, std::string>&) : data(new Data) {
    data->system = &system;
    data->platform = &platform;
    data->impl = NULL;
    build();
}

void Context::build() {
    Data& d = *data;
    const int n = d.system->getNumParticles();
    d.positions.assign(n, Vec3());
    d.velocities.assign(n, Vec3());
    d.forces.assign(n, Vec3());
    d.system->getDefaultPeriodicBoxVectors(d.box[0], d.box[1], d.box[2]);
    d.platformData.positions = &d.positions;
    d.platformData.velocities = &d.velocities;
    d.platformData.forces = &d.forces;
    d.platformData.periodicBoxVectors = d.box;
    d.platformData.energyParameterDerivatives = &d.derivatives;
    d.impl = new ContextImpl(*d.system, &d.platformData, d.platform);
    for (int i = 0; i < d.system->getNumForces(); i++)
        d.forceImpls.push_back(d.system->getForce(i).createImpl());
    for (size_t i = 0; i < d.forceImpls.size(); i++) {
        const std::map<std::string, double> defaults = d.forceImpls[i]->getDefaultParameters();
        for (std::map<std::string, double>::const_iterator p = defaults.begin(); p != defaults.end(); ++p)
            d.impl->setParameter(p->first, p->second);
    }
    for (size_t i = 0; i < d.forceImpls.size(); i++)
        d.forceImpls[i]->initialize(*d.impl);
}

void Context::destroy() {
    for (size_t i = 0; i < data->forceImpls.size(); i++)
        delete data->forceImpls[i];
    data->forceImpls.clear();
    delete data->impl;
    data->impl = NULL;
}

Context::~Context() {
    destroy();
    delete data;
}

const System& Context::getSystem() const { return *data->system; }
Platform& Context::getPlatform() { return *data->platform; }
void Context::setPositions(const std::vector<Vec3>& positions) {
    if ((int) positions.size() != data->system->getNumParticles())
        throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    data->positions = positions;
}
void Context::setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { data->box[0] = a; data->box[1] = b; data->box[2] = c; }
void Context::setParameter(const std::string& name, double value) {
    data->impl->getParameter(name);         /* throws for an unknown name, as OpenMM does */
    data->impl->setParameter(name, value);
}
double Context::getParameter(const std::string& name) const { return data->impl->getParameter(name); }
const std::map<std::string, double>& Context::getParameters() const { return data->impl->getParameters(); }

State Context::getState(int types, bool, int groups) const {
    Data& d = *data;
    State state;
    const bool includeForces = types&State::Forces, includeEnergy = types&State::Energy;
    if (includeForces || includeEnergy || (types&State::ParameterDerivatives)) {
        d.forces.assign(d.positions.size(), Vec3());
        d.derivatives.clear();
        double energy = 0;
        for (size_t i = 0; i < d.forceImpls.size(); i++)
            energy += d.forceImpls[i]->calcForcesAndEnergy(*d.impl, includeForces, includeEnergy, groups);
        state.energy = includeEnergy ? energy : 0.0;
        if (includeForces)
            state.forces = d.forces;
        if (types&State::ParameterDerivatives)
            state.derivatives = d.derivatives;
    }
    if (types&State::Positions)
        state.positions = d.positions;
    if (types&State::Parameters)
        state.parameters = d.impl->getParameters();
    state.box[0] = d.box[0]; state.box[1] = d.box[1]; state.box[2] = d.box[2];
    return state;
}

void Context::reinitialize(bool preserveState) {
    const std::vector<Vec3> positions = data->positions;
    const std::map<std::string, double> parameters = data->impl->getParameters();
    Vec3 box[3] = {data->box[0], data->box[1], data->box[2]};
    destroy();
    build();
    if (preserveState) {
        data->positions = positions;
        data->box[0] = box[0]; data->box[1] = box[1]; data->box[2] = box[2];
        for (std::map<std::string, double>::const_iterator p = parameters.begin(); p != parameters.end(); ++p)
            if (data->impl->getParameters().find(p->first) != data->impl->getParameters().end())
                data->impl->setParameter(p->first, p->second);
    }
}

/* ---- Force <-> Context ---- */
ForceImpl& Force::getImplInContext(Context& context) {
    for (size_t i = 0; i < context.data->forceImpls.size(); i++)
        if (&context.data->forceImpls[i]->getOwner() == this)
            return *context.data->forceImpls[i];
    throw OpenMMException("getImplInContext: This Force is not present in the Context");
}
const ForceImpl& Force::getImplInContext(const Context& context) const {
    return const_cast<Force*>(this)->getImplInContext(const_cast<Context&>(context));
}
ContextImpl& Force::getContextImpl(Context& context) { return *context.data->impl; }

void NonbondedForce::updateParametersInContext(Context& context) {
    dynamic_cast<PlainNonbondedForceImpl&>(getImplInContext(context)).updateParametersInContext(getContextImpl(context));
}

}
