"""cv::FileStorage-compatible XML output for params.xml / stage<i>.xml / cascade.xml.

The reference writes these files with cv::FileStorage (traincascade/lib/src/cascadeclassifier.cpp:
248-275,439-456).  To stay byte-compatible the same library does the formatting here: python cv2's
FileStorage (OpenCV 4.x) behind a small structural wrapper.  This is host-side glue only.
"""


class XmlWriter:
    def __init__(self, path):
        try:
            import cv2
        except ImportError as e:  # pragma: no cover
            raise RuntimeError("writing cascade XML needs python cv2 (cv::FileStorage): %s" % e)
        self.cv2 = cv2
        self.fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
        if not self.fs.isOpened():
            raise IOError("file %s can not be opened" % path)

    def start_map(self, name, type_name=None, flow=False):
        flags = self.cv2.FILE_NODE_MAP | (self.cv2.FILE_NODE_FLOW if flow else 0)
        if type_name:
            self.fs.startWriteStruct(name or "", flags, type_name)       # `fs << "{:" "type_id"`
        else:
            self.fs.startWriteStruct(name or "", flags)

    def end_map(self):
        self.fs.endWriteStruct()

    def start_seq(self, name):
        self.fs.startWriteStruct(name or "", self.cv2.FILE_NODE_SEQ)

    def end_seq(self):
        self.fs.endWriteStruct()

    def write(self, name, value):
        if isinstance(value, bool):
            value = int(value)
        self.fs.write(name or "", value)

    def write_flow_seq(self, name, values):
        """`fs << name << "[:" << v0 << v1 ... << "]"`"""
        self.fs.startWriteStruct(name or "", self.cv2.FILE_NODE_SEQ | self.cv2.FILE_NODE_FLOW)
        for v in values:
            self.fs.write("", v)
        self.fs.endWriteStruct()

    def comment(self, text):
        self.fs.writeComment(text)

    def release(self):
        self.fs.release()
