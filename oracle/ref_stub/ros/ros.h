// stand-in: cslam/Datatypes.h names this header; nothing of it is used by the translation units compiled here
