// TEST INFRASTRUCTURE — empty stand-in for the legacy <opencv/highgui.h> (included by the node, nothing of it is used)
#pragma once
