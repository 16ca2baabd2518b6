// TEST INFRASTRUCTURE — empty stand-in for the legacy <opencv/cvwimage.h> (included by the node, nothing of it is used)
#pragma once
