/* Cross-language check of the XML wire format: the reference's OWN serialization proxy (serialization/src/
 * SlicedNonbondedForceProxy.cpp, compiled unchanged, registered by its own OpenMMLabSerializationProxyRegistration.cpp) on the
 * stand-in XmlSerializer of plugin/mock_openmm.
 *     xml_cross write <file>      the force of the reference's serialization test -> XML
 *     xml_cross copy <in> <out>   XML -> SlicedNonbondedForce (the proxy's deserialize) -> XML
 * tests/test_serialization.py reads what `write` wrote with openmm-lab_b200/serialization.py, and sends what the Python side
 * wrote through `copy`. */
#include "SlicedNonbondedForce.h"
#include "openmm/serialization/XmlSerializer.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace OpenMM;
using namespace OpenMMLab;
using namespace std;

static SlicedNonbondedForce* testForce() {
    /* serialization/tests/TestSerializeSlicedNonbondedForce.cpp:25-58 */
    SlicedNonbondedForce* force = new SlicedNonbondedForce(3);
    force->setForceGroup(3);
    force->setName("custom name");
    force->setNonbondedMethod(SlicedNonbondedForce::CutoffPeriodic);
    force->setSwitchingDistance(1.5);
    force->setUseSwitchingFunction(true);
    force->setCutoffDistance(2.0);
    force->setEwaldErrorTolerance(1e-3);
    force->setReactionFieldDielectric(50.0);
    force->setUseDispersionCorrection(false);
    force->setExceptionsUsePeriodicBoundaryConditions(true);
    force->setIncludeDirectSpace(false);
    force->setPMEParameters(0.5, 3, 5, 7);
    force->setLJPMEParameters(0.8, 4, 6, 7);
    force->addParticle(1, 0.1, 0.01);
    force->addParticle(0.5, 0.2, 0.02);
    force->addParticle(-0.5, 0.3, 0.03);
    force->setParticleSubset(0, 1);
    force->setParticleSubset(1, 2);
    force->addException(0, 1, 2, 0.5, 0.1);
    force->addException(1, 2, 0.2, 0.4, 0.2);
    force->addGlobalParameter("scale1", 1.0);
    force->addGlobalParameter("scale2", 2.0);
    force->addParticleParameterOffset("scale1", 2, 1.5, 2.0, 2.5);
    force->addExceptionParameterOffset("scale2", 1, -0.1, -0.2, -0.3);
    force->addGlobalParameter("lambda", 0.5);
    force->addScalingParameter("lambda", 0, 1, true, true);
    force->addScalingParameter("lambda", 1, 1, false, true);
    force->addScalingParameterDerivative("lambda");
    return force;
}

int main(int argc, char* argv[]) {
    try {
        if (argc == 3 && strcmp(argv[1], "write") == 0) {
            SlicedNonbondedForce* force = testForce();
            ofstream out(argv[2]);
            XmlSerializer::serialize<SlicedNonbondedForce>(force, "Force", out);
            delete force;
            return 0;
        }
        if (argc == 4 && strcmp(argv[1], "copy") == 0) {
            ifstream in(argv[2]);
            SlicedNonbondedForce* force = XmlSerializer::deserialize<SlicedNonbondedForce>(in);
            ofstream out(argv[3]);
            XmlSerializer::serialize<SlicedNonbondedForce>(force, "Force", out);
            delete force;
            return 0;
        }
    }
    catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        return 1;
    }
    printf("usage: xml_cross write <file> | copy <in> <out>\n");
    return 2;
}
