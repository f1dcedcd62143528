from m3p2i_aip import extend_path

extend_path(__path__, "planners", "motion_planner")
