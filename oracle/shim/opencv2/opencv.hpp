// oracle/shim: included by the reference's Scancontext.hpp, nothing of it is used on the descriptor path (see README.md).
#pragma once
