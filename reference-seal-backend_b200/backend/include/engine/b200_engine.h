// b200_engine.h -- plugin entry of the B200 backend; same shape as R/include/engine/seal_engine.h.
#pragma once
#include "hebench/api_bridge/cpp/hebench.hpp"

#define HEBENCH_API_VERSION_NEEDED_MAJOR 0
#define HEBENCH_API_VERSION_NEEDED_MINOR 8
#define HEBENCH_API_VERSION_NEEDED_REVISION 0

class SEALEngine : public hebench::cpp::BaseEngine
{
public:
    HEBERROR_DECLARE_CLASS_NAME(SEALEngine)
    static SEALEngine *create();
    static void destroy(SEALEngine *p);
    ~SEALEngine() override;

protected:
    SEALEngine();
    void init() override;
};
