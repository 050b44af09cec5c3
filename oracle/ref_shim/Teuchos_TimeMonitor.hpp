#pragma once
#include "Teuchos_RCP.hpp"
namespace Teuchos { class TimeMonitor { public: TimeMonitor(Time &) {} }; }
