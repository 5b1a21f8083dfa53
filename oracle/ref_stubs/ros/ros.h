// TEST INFRASTRUCTURE ONLY: stands in for <ros/ros.h> when the reference MINCO header is compiled into oracle/_ref (the header includes it but uses nothing from it)
#pragma once
