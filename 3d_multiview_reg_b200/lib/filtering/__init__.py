from . import oanet

filtering_dict = {
    "oanet": oanet.OANet,
}
