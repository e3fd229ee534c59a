#pragma once
#include "BundleActivator.h"
