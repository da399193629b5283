// placeholder: filled in by the streaming kernel
#pragma once
