"""TEST-ONLY inert mock of rclpy (ROS2 python client); the ROS path is out of scope."""


class _Future:
    def result(self):
        return None


class _Client:
    srv_name = "mock"

    def wait_for_service(self, timeout_sec=None):
        return False

    def call_async(self, request):
        return _Future()


class _Logger:
    def info(self, msg):
        pass


class _Node:
    def create_client(self, srv_type, name):
        c = _Client()
        c.srv_name = name
        return c

    def get_logger(self):
        return _Logger()


def init(*args, **kwargs):
    pass


def create_node(name):
    return _Node()


def spin_until_future_complete(node, future):
    pass
