"""TEST-ONLY inert mock of the roboy_simulation_msgs ROS package."""
