// TEST INFRASTRUCTURE shim (console REPL unused by the reference: img_env.cpp:15-34).
#pragma once
