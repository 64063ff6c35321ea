PoseTracker = Nav2GoalAgent = None
