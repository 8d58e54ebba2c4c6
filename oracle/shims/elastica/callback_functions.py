"""TEST INFRASTRUCTURE (oracle) — CallBackBaseClass of PyElastica ([PE-recall])."""


class CallBackBaseClass:
    def __init__(self):
        pass

    def make_callback(self, system, time, current_step):
        pass
