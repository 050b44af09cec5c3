// TEST INFRASTRUCTURE ONLY (oracle): utils.h only needs the type name.
#pragma once
namespace Teuchos { class ParameterList {}; }
