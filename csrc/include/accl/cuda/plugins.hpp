// Device-API plugins (compute kernels that talk to the engine themselves).
#pragma once
