class _Srv:
    class Request:
        pass


class GymStep(_Srv):
    pass


class GymReset(_Srv):
    pass


class GymGoal(_Srv):
    pass
